"""The multi-GPU construction (kmercamel_b200/csrc/group.cuh behind kc_group_* / kc_init_multi).

CPU part (`-m "not gpu"`):
  * the job geometry every rank derives on its own (kc_group_plan, host-only) tiles the input and partitions the hash space;
  * world_size 2 over gloo: each rank extracts the k-mers of ITS slice, routes every item to the owner of its hash range as the
    plan says (numpy stand-in for the kernels, gloo all-to-all for the NVLink stores), resolves its range, and the OR of the
    ranks' first-occurrence bits must be exactly the bits of a single-process computation — the sharding itself is right;
  * the handle all-gather of sharded.attach.
GPU part (`-m gpu`): the real kernels and the real peer protocol.  Several ranks on ONE device (a device ordinal may repeat in
kc_init_multi) exercise every peer store, signal and wait of the N-GPU path on the one-GPU box the driver tests on, and two
PROCESSES sharing the GPU exercise the CUDA-IPC form that bench.py uses under torchrun; with >= 2 GPUs visible the same tests
spread the ranks over distinct devices.  The result must be byte-identical to kc_compute on one GPU.
"""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import kmercamel_b200 as kb
from kmercamel_b200 import sharded, synth

MULT = np.uint64(0x9E3779B97F4A7C15)


def window_kmers(seq: np.ndarray, k: int, complements: bool):
    """(canonical k-mer as uint64, window END position) of every valid window, k <= 31 (plain numpy, test only)."""
    lut = np.full(256, 4, dtype=np.uint8)
    for i, c in enumerate(b"ACGT"):
        lut[c] = i
        lut[c + 32] = i
    codes = lut[seq]
    n = len(seq)
    if n < k:
        return np.zeros(0, np.uint64), np.zeros(0, np.int64)
    bad = (codes > 3).astype(np.int64)
    csum = np.concatenate([[0], np.cumsum(bad)])
    ends = np.arange(k - 1, n)
    valid = (csum[ends + 1] - csum[ends + 1 - k]) == 0
    fwd = np.zeros(n - k + 1, dtype=np.uint64)
    rc = np.zeros(n - k + 1, dtype=np.uint64)
    c64 = (codes & 3).astype(np.uint64)
    for j in range(k):
        col = c64[j:j + n - k + 1]
        fwd |= col << np.uint64(2 * (k - 1 - j))
        rc |= (np.uint64(3) - col) << np.uint64(2 * j)
    canon = np.minimum(fwd, rc) if complements else fwd
    return canon[valid], ends[valid]


def first_occurrence_flags(seq, k, complements, min_frequency):
    keys, pos = window_kmers(seq, k, complements)
    words = np.zeros((len(seq) + 31) // 32 + 1, dtype=np.uint32)
    if len(keys) == 0:
        return words, 0
    order = np.lexsort((pos, keys))
    ks, ps = keys[order], pos[order]
    head = np.concatenate([[True], ks[1:] != ks[:-1]])
    starts = np.flatnonzero(head)
    counts = np.diff(np.concatenate([starts, [len(ks)]]))
    keep = counts >= min_frequency
    first = ps[starts[keep]]
    np.bitwise_or.at(words, first >> 5, (np.uint32(1) << (first & 31).astype(np.uint32)))
    return words, int(keep.sum())


def make_input(seed):
    recs = synth.random_genome_records(3, 700, seed)
    recs.append(recs[0][100:400].copy())          # duplicated region
    recs.append(np.frombuffer(b"ACGTNNACGTACGTTTGACCA", dtype=np.uint8).copy())
    reads = synth.reads_from_genome(500, 6.0, 60, 0.02, seed)
    seq, _, _ = synth.frame_records(recs + list(reads))
    return seq


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]



# ---- CPU: geometry -------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("world", [1, 2, 3, 4, 5, 8, 16])
def test_group_plan_invariants(world):
    for k in (15, 31, 63, 127):
        for n_bytes in (0, 1000, 70_000, 8192 * 5 + 3, 3_000_001, 50_000_050, 400_000_400, 3_100_000_024):
            plans = sharded.check_plans(world, k, n_bytes)
            g = int(kb.load_library().kc_shard_granule(k))
            assert all(p["pos_begin"] % g == 0 and (p["pos_end"] % g == 0 or p["pos_end"] == n_bytes) for p in plans)
            if plans[0]["fixed_slots"]:
                # a sub-slot holds the expected share of a slice with room to spare, and stays 128-byte aligned for the positions
                slice_max = max(p["pos_end"] - p["pos_begin"] for p in plans)
                assert plans[0]["cap_sub"] % 32 == 0 and plans[0]["cap_sub"] > slice_max / plans[0]["n_digits"]
                assert all(p["digit_end"] > p["digit_begin"] for p in plans)
    assert not sharded.group_plan(2, 0, 31, 50_000)["fixed_slots"]       # too small: the exact construction
    assert sharded.group_plan(8, 3, 31, 400_000_400)["fixed_slots"]


# ---- CPU: the sharding is a correct decomposition (world 2 over gloo, numpy stand-in for the kernels) -------------------
def _sharded_flags_cpu(rank, world, seq, k, complements, z):
    plan = sharded.group_plan(world, rank, k, len(seq))
    plans = [sharded.group_plan(world, r, k, len(seq)) for r in range(world)]
    keys, pos = window_kmers(seq, k, complements)
    mine = (pos >= plan["pos_begin"]) & (pos < plan["pos_end"])
    keys, pos = keys[mine], pos[mine]
    bits = int(plan["n_digits"]).bit_length() - 1
    digit = ((keys * MULT) >> np.uint64(64 - bits)).astype(np.int64)       # level-0 digit of the scrambled word
    owner = np.searchsorted(np.array([p["digit_end"] for p in plans]), digit, side="right")
    order = np.argsort(owner, kind="stable")
    send = np.bincount(owner, minlength=world).astype(np.int64)
    s = torch.from_numpy(send)
    r = torch.empty_like(s)
    dist.all_to_all_single(r, s)
    recv = r.numpy()
    rk = torch.empty(int(recv.sum()), dtype=torch.int64)
    rp = torch.empty(int(recv.sum()), dtype=torch.int64)
    dist.all_to_all_single(rk, torch.from_numpy(keys[order].astype(np.int64)), recv.tolist(), send.tolist())
    dist.all_to_all_single(rp, torch.from_numpy(pos[order].astype(np.int64)), recv.tolist(), send.tolist())
    ks, ps = rk.numpy().astype(np.uint64), rp.numpy()
    words = np.zeros((len(seq) + 31) // 32 + 1, dtype=np.uint32)
    kept = 0
    if len(ks):
        o = np.lexsort((ps, ks))
        ks, ps = ks[o], ps[o]
        starts = np.flatnonzero(np.concatenate([[True], ks[1:] != ks[:-1]]))
        counts = np.diff(np.concatenate([starts, [len(ks)]]))
        keep = counts >= z
        first = ps[starts[keep]]
        np.bitwise_or.at(words, first >> 5, (np.uint32(1) << (first & 31).astype(np.uint32)))
        kept = int(keep.sum())
    t = torch.from_numpy(words.view(np.int32).copy())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)                               # disjoint bits: SUM == OR
    tot = torch.tensor([kept, int(send.sum()) - int(send[rank])], dtype=torch.int64)
    dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    return t.numpy().view(np.uint32), int(tot[0]), int(tot[1])


def _worker(rank, world, port, k, complements, z, seed, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        seq = make_input(seed)
        seq = np.concatenate([seq] + [synth.frame_records(synth.random_genome_records(2, 40_000, seed + 1))[0]])
        handles = sharded.gather_handles(np.full(64, rank + 1, dtype=np.uint8), world)      # the set-up plumbing of sharded.attach
        flags, kept, moved = _sharded_flags_cpu(rank, world, seq, k, complements, z)
        if rank == 0:
            want, want_kept = first_occurrence_flags(seq, k, complements, z)
            out.put((bool(np.array_equal(flags, want)), kept == want_kept, moved, handles.tolist() == [1] * 64 + [2] * 64))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("k,complements,z", [(11, True, 1), (21, False, 1), (15, True, 2), (31, True, 3)])
def test_sharding_two_ranks_gloo(k, complements, z):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, k, complements, z, 7 + k, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    flags_ok, kept_ok, moved, handles_ok = out.get(timeout=10)
    assert flags_ok and kept_ok and handles_ok
    assert moved > 0                          # the exchange really moved items between the two ranks


# ---- GPU: several ranks in one process ----------------------------------------------------------------------------------------
def _devices(n_ranks):
    n = torch.cuda.device_count()
    return [r % n for r in range(n_ranks)]


def _inputs():
    g = synth.frame_records(synth.random_genome_records(5, 600_000, 41))[0]
    rep = synth.frame_records(synth.human_like_genome(2_000_000, 43))[0]
    reads = synth.frame_reads(synth.reads_chunks(60_000, 12.0, 150, 0.01, 44, chunk_reads=2000), 150)
    return {"genome": g, "repeats": rep, "reads": reads, "twice": np.concatenate([g, g[: len(g) // 2]])}


@pytest.mark.gpu
@pytest.mark.parametrize("n_ranks", [1, 2, 3, 4, 8])
def test_group_in_process_matches_single_gpu(ctx, n_ranks):
    inputs = _inputs()
    grp = kb.Group(_devices(n_ranks))
    try:
        for name, k, compl, z in [("genome", 31, True, 1), ("repeats", 31, True, 1), ("twice", 31, False, 1), ("genome", 63, True, 1),
                                  ("repeats", 127, False, 1), ("reads", 31, True, 2), ("reads", 21, True, 1), ("genome", 31, True, 1)]:
            seq = inputs[name]
            want = ctx.compute(seq, k=k, complements=compl, min_frequency=z)
            got = grp.compute(seq, k=k, complements=compl, min_frequency=z)
            assert (got.n_kmers, got.length, got.n_nodes) == (want.n_kmers, want.length, want.n_nodes), (name, k, z)
            assert got.ms == want.ms, (name, k, z)
        runs, sig = grp.stat("fast_runs"), grp.stat("sig_runs")
        assert min(sig) >= 5 and len(set(sig)) == 1            # k >= 26 without -z: signature buckets on every rank
        assert len(set(runs)) == 1 and min(grp.stat("sig_fallbacks")) == 0
        # the same jobs through the fixed-slot construction
        grp.set_option("sig_set", 0)
        for name, k, compl, z in [("genome", 31, True, 1), ("repeats", 127, False, 1)]:
            seq = inputs[name]
            assert grp.compute(seq, k=k, complements=compl, min_frequency=z).ms == ctx.compute(seq, k=k, complements=compl, min_frequency=z).ms
        assert min(grp.stat("fast_runs")) >= min(runs) + 2
    finally:
        grp.close()


@pytest.mark.gpu
def test_group_overflow_falls_back_on_every_rank(ctx):
    """Slots planned without slack overflow; every rank sees the status word and all switch to the exact construction together."""
    seq = _inputs()["repeats"]
    want = ctx.compute(seq, k=31)
    grp = kb.Group(_devices(3))
    try:
        grp.set_option("fast_heuristics", 0)
        grp.set_option("sig_load_pct", 400)        # signature buckets at four times their capacity: every rank falls back together
        grp.set_option("fast_sigmas", 0)           # ... to fixed slots without slack, which overflow as well
        for _ in range(2):
            got = grp.compute(seq, k=31)
            assert got.ms == want.ms
        assert min(grp.stat("sig_fallbacks")) >= 2 and min(grp.stat("fast_fallbacks")) >= 2
        grp.set_option("sig_load_pct", 0)
        assert grp.compute(seq, k=31).ms == want.ms and min(grp.stat("sig_runs")) >= 1
        grp.set_option("sig_set", 0)
        grp.set_option("fast_sigmas", 8)
        assert grp.compute(seq, k=31).ms == want.ms and min(grp.stat("fast_runs")) >= 1
    finally:
        grp.close()


@pytest.mark.gpu
def test_group_small_and_empty_inputs(ctx):
    grp = kb.Group(_devices(2))
    try:
        small = synth.frame_records(synth.random_genome_records(3, 5_000, 3))[0]
        assert grp.compute(small, k=31).ms == ctx.compute(small, k=31).ms            # below the fixed-slot plan: exact path
        with pytest.raises(kb.KcError) as e:
            grp.compute(np.frombuffer(b"ACGT\nNNNN\n", dtype=np.uint8), k=31)
        assert e.value.code == -4                                                     # no k-mers: KC_ERR_EMPTY on the group too
        assert grp.compute(small, k=15).ms == ctx.compute(small, k=15).ms            # and the group is still usable
        with pytest.raises(kb.KcError):
            grp.compute(small, k=31, min_frequency=300)
    finally:
        grp.close()


# ---- GPU: one process per rank (CUDA IPC heaps), launched as the driver launches bench.py ------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4])
def test_group_torchrun_processes_match_single_gpu(tmp_path, world):
    out = tmp_path / "result.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(os.path.dirname(__file__), "sharded_worker.py"), str(out)]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    res = json.load(open(out))
    assert res["world"] == world and res["cases"] >= 5 and res["all_identical"], res

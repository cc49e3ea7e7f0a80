"""Host-side logic of the hash-range sharded path (kmercamel_b200/sharded.py).

CPU part (`-m "not gpu"`): the orchestration runs with world_size 2 over gloo, with a numpy stand-in for the two GPU
halves (partition / resolve), and must produce exactly the first-occurrence flags of a single-process computation.
GPU part (`-m gpu`): world_size 1 through the real library halves must reproduce kc_compute byte for byte.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

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


class CpuStandInOps:
    """numpy stand-in for GpuOps: same interface, same data layout (items grouped by level-0 digit)."""

    def __init__(self, seq: np.ndarray):
        self.seq = seq
        self.n_bytes = len(seq)
        self.flags = torch.zeros((self.n_bytes + 31) // 32 + 1, dtype=torch.int32)

    def granule(self, k):
        return 64

    def partition(self, b, e, *, k, complements):
        keys, pos = window_kmers(self.seq, k, complements)
        sel = (pos >= b) & (pos < e)
        keys, pos = keys[sel], pos[sel]
        digit = ((keys * MULT) >> np.uint64(56)).astype(np.int64)
        order = np.argsort(digit, kind="stable")
        self._send_k = torch.from_numpy(keys[order].astype(np.int64))
        self._send_p = torch.from_numpy(pos[order].astype(np.int32))
        return np.bincount(digit, minlength=256).astype(np.int64), len(keys)

    def exchange_items(self, comm, send, recv):
        n_recv = int(recv.sum())
        keys = torch.empty(n_recv, dtype=torch.int64)
        pos = torch.empty(n_recv, dtype=torch.int32)
        comm.all_to_all(keys, self._send_k, recv, send, 1)
        comm.all_to_all(pos, self._send_p, recv, send, 1)
        return keys, pos

    def resolve(self, keys, pos, n, *, k, complements, min_frequency):
        self.flags.zero_()
        ks = keys.numpy()[:n].astype(np.uint64)
        ps = pos.numpy()[:n].astype(np.int64)
        if n == 0:
            return 0
        order = np.lexsort((ps, ks))
        ks, ps = ks[order], ps[order]
        head = np.concatenate([[True], ks[1:] != ks[:-1]])
        starts = np.flatnonzero(head)
        counts = np.diff(np.concatenate([starts, [n]]))
        keep = counts >= min_frequency
        first = ps[starts[keep]]
        words = np.zeros(self.flags.numel(), dtype=np.uint32)
        np.bitwise_or.at(words, first >> 5, (np.uint32(1) << (first & 31).astype(np.uint32)))
        self.flags.copy_(torch.from_numpy(words.view(np.int32)))
        return int(keep.sum())

    def reduce_flags(self, comm, all_ranks: bool = False):
        if all_ranks:
            comm.all_reduce_sum(self.flags)
        else:
            comm.reduce_sum(self.flags, 0)

    def finish(self, n_kept, *, k, complements, slice=None):
        return self.flags.numpy().view(np.uint32).copy()


def make_input(seed):
    recs = synth.random_genome_records(3, 700, seed)
    recs.append(recs[0][100:400].copy())          # duplicated region
    recs.append(np.frombuffer(b"ACGTNNACGTACGTTTGACCA", dtype=np.uint8).copy())
    reads = synth.reads_from_genome(500, 6.0, 60, 0.02, seed)
    seq, _, _ = synth.frame_records(recs + list(reads))
    return seq


def _worker(rank, world, port, k, complements, z, seed, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        seq = make_input(seed)
        ops = CpuStandInOps(seq)
        comm = sharded.TorchComm(torch.device("cpu"))
        r = sharded.sharded_compute(ops, comm, len(seq), k=k, complements=complements, min_frequency=z)
        if rank == 0:
            want, want_kept = first_occurrence_flags(seq, k, complements, z)
            keys, _ = window_kmers(seq, k, complements)
            out.put((bool(np.array_equal(r.result, want)), r.n_kept == want_kept, r.n_occurrences == len(keys), r.items_sent, r.items_received))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("k,complements,z", [(11, True, 1), (21, False, 1), (15, True, 2), (31, True, 3)])
def test_sharded_two_ranks_gloo(k, complements, z):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, k, complements, z, 7 + k, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    flags_ok, kept_ok, occ_ok, sent, received = out.get(timeout=10)
    assert flags_ok and kept_ok and occ_ok
    assert sent > 0 and received > 0          # the exchange really moved items between the two ranks


def test_plan_slices_and_owners():
    for n_bytes, world, g in [(50_000_050, 8, 8192), (1000, 4, 8192), (8192 * 5 + 3, 3, 8192), (0, 2, 64), (777, 1, 64)]:
        sl = sharded.plan_slices(n_bytes, world, g)
        assert len(sl) == world and sl[0][0] == 0 and sl[-1][1] == n_bytes
        for (b, e), (b2, _) in zip(sl, sl[1:]):
            assert e == b2 and b <= e
        assert all(b % g == 0 for b, _ in sl) and all(e % g == 0 or e == n_bytes for _, e in sl)
    for world in (1, 2, 3, 4, 8):
        owners = [sharded.owner_of_digit(d, world) for d in range(256)]
        assert owners == sorted(owners) and set(owners) == set(range(world))
        counts = np.arange(256)
        oc = sharded.owner_counts(counts, world)
        assert oc.sum() == counts.sum() and all(oc[r] == sum(d for d in range(256) if owners[d] == r) for r in range(world))


def test_single_rank_stand_in_matches_flags():
    seq = make_input(3)
    ops = CpuStandInOps(seq)
    comm = sharded.TorchComm(torch.device("cpu"))
    r = sharded.sharded_compute(ops, comm, len(seq), k=13, complements=True, min_frequency=1)
    want, kept = first_occurrence_flags(seq, 13, True, 1)
    assert np.array_equal(r.result, want) and r.n_kept == kept


@pytest.mark.gpu
@pytest.mark.parametrize("k,complements,z", [(31, True, 1), (31, True, 2), (63, False, 1), (127, True, 1), (9, True, 1)])
def test_sharded_world1_matches_kc_compute(ctx, k, complements, z):
    recs = synth.random_genome_records(6, 40_000, 5)
    recs.append(recs[1][5000:9000].copy())
    reads = synth.reads_from_genome(20_000, 8.0, 150, 0.01, seed=k)
    seq, _, _ = synth.frame_records(recs + list(reads))
    want = ctx.compute(seq, k=k, complements=complements, min_frequency=z)
    d = torch.from_numpy(seq).cuda()
    torch.cuda.synchronize()
    ops = sharded.GpuOps(ctx, d)
    comm = sharded.TorchComm(d.device)
    r = sharded.sharded_compute(ops, comm, d.numel(), k=k, complements=complements, min_frequency=z)
    assert r.n_kept == want.n_kmers and r.result.length == want.length
    assert ctx.copy_to_host(r.result.ms_ptr, r.result.length) == want.ms


@pytest.mark.gpu
@pytest.mark.parametrize("stream_mode", ["private", "torch"])
def test_sharded_p2p_world1_matches_kc_compute(stream_mode):
    """The fused partition + exchange path with a single rank (its own buffers stand in for the peers'), on a private
    library stream and on torch's current stream (then no host synchronisation separates torch's work from the kernels)."""
    import kmercamel_b200 as kb
    c = kb.Context(0, None if stream_mode == "private" else torch.cuda.current_stream().cuda_stream)
    try:
        recs = synth.random_genome_records(6, 40_000, 11)
        recs.append(recs[2][100:3000].copy())
        seq, _, _ = synth.frame_records(recs)
        d = torch.from_numpy(seq).cuda()
        ops = sharded.GpuOps(c, d)
        comm = sharded.TorchComm(d.device)
        ops.setup_p2p(comm, 31)
        for z in (1, 2):
            want = c.compute(seq, k=31, min_frequency=z)
            r = sharded.sharded_compute_p2p(ops, comm, d.numel(), k=31, min_frequency=z)
            assert r.n_kept == want.n_kmers and r.result.length == want.length
            assert c.copy_to_host(r.result.ms_ptr, r.result.length) == want.ms
    finally:
        c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("k,n_slices", [(31, 3), (63, 8), (15, 2)])
def test_sliced_emission_tiles_the_superstring(k, n_slices):
    """kc_compute_from_flags_slice: the slices of indices 0..n-1 (what the ranks of a multi-GPU job emit) concatenate to exactly
    the superstring of the unsliced call, for long runs (chunked 16-byte emission) and for thousands of short ones."""
    import kmercamel_b200 as kb
    c = kb.Context(0, torch.cuda.current_stream().cuda_stream)
    try:
        recs = synth.random_genome_records(3, 200_000, 31) + list(synth.reads_from_genome(30_000, 6.0, 150, 0.01, seed=k))
        seq, _, _ = synth.frame_records(recs)
        d = torch.from_numpy(seq).cuda()
        ops = sharded.GpuOps(c, d)
        comm = sharded.TorchComm(d.device)
        ops.setup_p2p(comm, k)
        whole = sharded.sharded_compute_p2p(ops, comm, d.numel(), k=k)
        want = c.copy_to_host(whole.result.ms_ptr, whole.result.length)
        assert want == kb.Context(0).compute(seq, k=k).ms
        parts, at = [], 0
        for i in range(n_slices):
            r = c.compute_from_flags(d.data_ptr(), d.numel(), ops.flags.data_ptr(), whole.n_kept, k=k, slice=(i, n_slices))
            assert r.length == len(want) and r.slice_begin == at and (r.slice_begin % 16 == 0)
            parts.append(c.copy_to_host(r.ms_ptr, r.slice_len))
            at += r.slice_len
        assert at == len(want) and b"".join(parts) == want
        sliced = sharded.sharded_compute_p2p(ops, comm, d.numel(), k=k, slice_output=True)     # world 1: one slice = everything
        assert (sliced.result.slice_begin, sliced.result.slice_len) == (0, len(want))
        assert c.copy_to_host(sliced.result.ms_ptr, sliced.result.slice_len) == want
    finally:
        c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("k,z", [(31, 1), (31, 2), (47, 1), (100, 2)])
def test_sharded_p2p_fixed_slot_resolve_with_duplicates(k, z):
    """Owner-side fixed-slot levels (kc_kmerset_resolve_fast) on inputs where every k-mer occurs twice or three times, with the
    data set changing between jobs (stale receive-buffer contents must never leak into a result)."""
    import kmercamel_b200 as kb
    c = kb.Context(0, torch.cuda.current_stream().cuda_stream)
    one = kb.Context(0)
    try:
        comm = sharded.TorchComm(torch.device("cuda", 0))
        a = synth.frame_records(synth.random_genome_records(4, 500_000, 21))[0]
        b = synth.frame_records(synth.random_genome_records(4, 500_000, 22))[0]
        first = True
        for parts in ((a, b), (b, b), (a, a[:700_000], b, a), (a, b)):
            seq = np.concatenate(parts)
            d = torch.from_numpy(seq).cuda()
            ops = sharded.GpuOps(c, d)
            if first:
                ops.setup_p2p(comm, k, slack=2.0)
                first = False
            try:
                want = one.compute(seq, k=k, min_frequency=z)
            except kb.api.KcError as e:      # all k-mers distinct and z = 2: nothing is kept
                assert e.code == -4
                with pytest.raises(kb.api.KcError):
                    sharded.sharded_compute_p2p(ops, comm, d.numel(), k=k, min_frequency=z)
                continue
            r = sharded.sharded_compute_p2p(ops, comm, d.numel(), k=k, min_frequency=z)
            assert r.n_kept == want.n_kmers and r.result.length == want.length
            assert c.copy_to_host(r.result.ms_ptr, r.result.length) == want.ms
        assert c.stat("fast_runs") >= 3
    finally:
        c.close()
        one.close()

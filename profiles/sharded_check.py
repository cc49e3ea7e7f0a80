"""Multi-GPU parity check + phase times (diagnosis, not a bench value):  the sharded job (fused partition + exchange over
peer memory) must produce, on rank 0, exactly the superstring that one GPU computes from the same sequence.
torchrun --nproc-per-node N profiles/sharded_check.py"""
import hashlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
import kmercamel_b200 as kb
from kmercamel_b200 import sharded, synth

rank = int(os.environ.get("RANK", 0)); lr = int(os.environ.get("LOCAL_RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
dev = torch.device("cuda", lr)
K = 31
ctx = kb.Context(lr, torch.cuda.current_stream().cuda_stream)
comm = sharded.TorchComm(dev)
# A: every rank contributes a different 50 Mbp part (all k-mers distinct); B: the same part on every rank (every k-mer `world` times)
for name, seed, zs in (("A", 12345 + rank, (1,)), ("B", 777, (1, 2))):
    part, _, _ = synth.frame_records(synth.random_genome_records(50, 1_000_000, seed))
    own = torch.from_numpy(part).to(dev)
    full = torch.empty(world * own.numel(), dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(full, own)
    ops = sharded.GpuOps(ctx, full)
    if name == "A":
        ops.setup_p2p(comm, K)
    for z in zs:
        r = sharded.sharded_compute_p2p(ops, comm, full.numel(), k=K, min_frequency=z)
        if rank == 0:
            got = ctx.copy_to_host(r.result.ms_ptr, r.result.length)
            ctx1 = kb.Context(lr, torch.cuda.current_stream().cuda_stream)
            one = ctx1.compute_device(full.data_ptr(), full.numel(), k=K, min_frequency=z)
            want = ctx1.copy_to_host(one.ms_ptr, one.length)
            ok = got == want and r.n_kept == one.n_kmers
            print(f"{name} z={z}: sharded kept={r.n_kept} len={len(got)} | single kept={one.n_kmers} len={one.length} | identical={ok} md5={hashlib.md5(got).hexdigest()}", flush=True)
            ctx1.close()
        dist.barrier()
        # sliced output: every rank emits its part; rank 0 checks that the parts tile the single-GPU superstring
        rs = sharded.sharded_compute_p2p(ops, comm, full.numel(), k=K, min_frequency=z, slice_output=True)
        mine = ctx.copy_to_host(rs.result.ms_ptr, rs.result.slice_len)
        parts = [None] * world
        dist.all_gather_object(parts, (rs.result.slice_begin, hashlib.md5(mine).hexdigest(), len(mine)))
        if rank == 0:
            ok = sum(p[2] for p in parts) == len(want) and all(hashlib.md5(want[b:b + n]).hexdigest() == h for b, h, n in parts)
            print(f"{name} z={z}: sliced output over {world} ranks tiles the single-GPU superstring: {ok}", flush=True)
        dist.barrier()
print(f"rank {rank}: fast_runs={ctx.stat('fast_runs')} fast_fallbacks={ctx.stat('fast_fallbacks')}", flush=True)
def T():
    torch.cuda.synchronize(); return time.perf_counter()
for it in range(5):
    dist.barrier()
    t = [T()]
    b, e = sharded.plan_slices(full.numel(), world, ops.granule(K))[rank]
    counts = ops.p2p_hist(b, e, k=K, complements=True); t.append(T())
    allc = comm.all_gather_counts(counts); t.append(T())
    ops.p2p_scatter(b, e, allc, k=K, complements=True); t.append(T())
    comm.barrier(); t.append(T())
    kept, owned = ops.p2p_resolve(allc, k=K, complements=True, min_frequency=1); t.append(T())
    ops.reduce_flags(comm); t.append(T())
    tot = comm.sum_scalars([kept, int(counts.sum())]); t.append(T())
    res = ops.finish(int(tot[0]), k=K, complements=True) if rank == 0 else None; t.append(T())
    names = ["hist", "gather_counts", "scatter_p2p", "barrier", "resolve", "reduce_flags", "sum_scalars", "finish"]
    if it >= 3 and rank in (0, world - 1):
        print(f"rank {rank} it {it}: " + "  ".join(f"{nm} {1000*(t[i+1]-t[i]):.3f}" for i, nm in enumerate(names)) + f"  total {1000*(t[-1]-t[0]):.3f} ms", flush=True)
dist.destroy_process_group()

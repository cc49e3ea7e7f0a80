"""Per-phase wall time of the sharded step (synchronising after every phase) — diagnosis only, not a bench value.
torchrun --nproc-per-node N profiles/sharded_phases.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
import kmercamel_b200 as kb
from kmercamel_b200 import sharded, synth

rank = int(os.environ.get("RANK", 0)); lr = int(os.environ.get("LOCAL_RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
dev = torch.device("cuda", lr)
recs = synth.random_genome_records(50, 1_000_000, 12345 + rank)
part, _, _ = synth.frame_records(recs)
own = torch.from_numpy(part).to(dev)
full = torch.empty(world * own.numel(), dtype=torch.uint8, device=dev)
if world > 1:
    dist.all_gather_into_tensor(full, own)
else:
    full.copy_(own)
ctx = kb.Context(lr, torch.cuda.current_stream().cuda_stream)
comm = sharded.TorchComm(dev); ops = sharded.GpuOps(ctx, full)
K = 31
def T():
    torch.cuda.synchronize(); return time.perf_counter()
P2P = len(sys.argv) > 1 and sys.argv[1] == "p2p"
if P2P:
    ops.setup_p2p(comm, K)
for it in range(6 if not P2P else 0):
    if world > 1: dist.barrier()
    t = [T()]
    b, e = sharded.plan_slices(full.numel(), world, ops.granule(K))[rank]
    dc, n = ops.partition(b, e, k=K, complements=True); t.append(T())
    send = sharded.owner_counts(dc, world); recv = comm.exchange_counts(send); t.append(T())
    keys, pos = ops.exchange_items(comm, send, recv); t.append(T())
    kept = ops.resolve(keys, pos, int(recv.sum()), k=K, complements=True, min_frequency=1); t.append(T())
    ops.reduce_flags(comm); t.append(T())
    tot = comm.sum_scalars([kept, n]); t.append(T())
    res = ops.finish(int(tot[0]), k=K, complements=True) if rank == 0 else None; t.append(T())
    names = ["partition", "counts", "all_to_all", "resolve", "reduce_flags", "sum_scalars", "finish"]
    if it >= 3:
        print(f"rank {rank} it {it}: " + "  ".join(f"{nm} {1000*(t[i+1]-t[i]):.3f}" for i, nm in enumerate(names)) + f"  total {1000*(t[-1]-t[0]):.3f} ms", flush=True)
for it in range(6 if P2P else 0):
    if world > 1: dist.barrier()
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
    if it >= 3:
        print(f"rank {rank} it {it}: " + "  ".join(f"{nm} {1000*(t[i+1]-t[i]):.3f}" for i, nm in enumerate(names)) + f"  total {1000*(t[-1]-t[0]):.3f} ms" + (f" kmers {int(tot[0])} len {res.length}" if res else ""), flush=True)
if world > 1:
    dist.destroy_process_group()

"""BASELINE.json configs[4] (3.1 Gbp, k = 31) on N GPUs: the sharded job against the single-GPU result of the same sequence
(rank 0 computes it first), plus timings.  One JSON line on stdout (rank 0).  Diagnosis / reporting only.
torchrun --nproc-per-node N profiles/sharded_big.py [--bases 3.1e9]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
import kmercamel_b200 as kb
from kmercamel_b200 import sharded

rank = int(os.environ.get("RANK", 0)); lr = int(os.environ.get("LOCAL_RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
dev = torch.device("cuda", lr)
K = 31
bases = float(sys.argv[sys.argv.index("--bases") + 1]) if "--bases" in sys.argv else 3.1e9
n_rec, rec_len = 31, int(bases / 31)
LUT = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
g = torch.Generator(device=dev).manual_seed(3100)          # same generator as profiles/run_big_configs.py cfg5: identical on every rank
full = torch.empty((n_rec, rec_len + 1), dtype=torch.uint8, device=dev)
for r in range(n_rec):
    full[r, :rec_len] = LUT[torch.randint(0, 4, (rec_len,), generator=g, device=dev, dtype=torch.uint8).long()]
full[:, rec_len] = 10
full = full.flatten()
torch.cuda.synchronize()
out = {"config": "BASELINE configs[4]: %d x %d bp uniform genome, k=31" % (n_rec, rec_len), "n_gpus": world, "n_bytes": int(full.numel())}
want = None
if rank == 0:
    one = kb.Context(lr, torch.cuda.current_stream().cuda_stream)
    r1 = one.compute_device(full.data_ptr(), full.numel(), k=K)
    r1 = one.compute_device(full.data_ptr(), full.numel(), k=K)
    want = np.empty(r1.length, dtype=np.uint8)
    one._check(one._lib.kc_copy_to_host(one._h, want.ctypes.data, r1.ms_ptr, r1.length))
    out["single_gpu"] = {"device_ms": r1.times_ms["total"], "n_kmers": r1.n_kmers, "length": r1.length, "nodes": r1.n_nodes}
    one.close()
    torch.cuda.empty_cache()
dist.barrier()
ctx = kb.Context(lr, torch.cuda.current_stream().cuda_stream)
comm = sharded.TorchComm(dev)
ops = sharded.GpuOps(ctx, full)
ops.setup_p2p(comm, K)
times = []
for it in range(4):
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = sharded.sharded_compute_p2p(ops, comm, full.numel(), k=K)
    torch.cuda.synchronize(); dist.barrier()
    times.append((time.perf_counter() - t0) * 1e3)
if rank == 0:
    got = np.empty(r.result.length, dtype=np.uint8)
    ctx._check(ctx._lib.kc_copy_to_host(ctx._h, got.ctypes.data, r.result.ms_ptr, r.result.length))
    out["sharded"] = {"wall_ms_per_job": times[1:], "n_kmers": r.n_kept, "length": r.result.length, "nodes": r.result.n_nodes,
                      "fast_runs": ctx.stat("fast_runs"), "fast_fallbacks": ctx.stat("fast_fallbacks"),
                      "kmers_per_s": r.n_kept / (min(times[1:]) * 1e-3)}
    out["identical_to_single_gpu"] = bool(r.n_kept == out["single_gpu"]["n_kmers"] and got.shape == want.shape and np.array_equal(got, want))
    out["ones_equal_kmers"] = int(np.count_nonzero(got <= 90)) == r.n_kept
    print(json.dumps(out), flush=True)
dist.barrier()
dist.destroy_process_group()

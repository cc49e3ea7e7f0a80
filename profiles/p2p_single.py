"""world = 1 exercise of the fused partition + exchange path (peer table = own buffers): results must equal kc_compute."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import kmercamel_b200 as kb
from kmercamel_b200 import sharded, synth
dev = torch.device("cuda", 0)
K = 31
ctx = kb.Context(0, torch.cuda.current_stream().cuda_stream)
ctx1 = kb.Context(0, torch.cuda.current_stream().cuda_stream)
comm = sharded.TorchComm(dev)
first = True
def dataset(name):
    a = synth.frame_records(synth.random_genome_records(20, 1_000_000, 1))[0]
    b = synth.frame_records(synth.random_genome_records(20, 1_000_000, 2))[0]
    if name == "A": return np.concatenate([a, b])
    if name == "B": return np.concatenate([b, b])
    if name == "C": return np.concatenate([a, a[:7_000_000], b[:13_000_000], a])
for name, z in (("A", 1), ("B", 1), ("B", 2), ("C", 1), ("A", 1), ("C", 2)):
    full = torch.from_numpy(dataset(name)).to(dev)
    ops = sharded.GpuOps(ctx, full)
    if first:
        ops.setup_p2p(comm, K, slack=1.5); first = False
    try:
        r = sharded.sharded_compute_p2p(ops, comm, full.numel(), k=K, min_frequency=z)
        got = ctx.copy_to_host(r.result.ms_ptr, r.result.length)
        one = ctx1.compute_device(full.data_ptr(), full.numel(), k=K, min_frequency=z)
        want = ctx1.copy_to_host(one.ms_ptr, one.length)
        print(f"{name} z={z}: p2p kept={r.n_kept} len={len(got)} | single kept={one.n_kmers} len={one.length} | identical={got == want}", flush=True)
    except Exception as e:
        print(f"{name} z={z}: FAILED {e}", flush=True)
        break
print("fast_runs", ctx.stat("fast_runs"), "fallbacks", ctx.stat("fast_fallbacks"))

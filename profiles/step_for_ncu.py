"""One warm-up + STEPS passes of the bench workload (configs[1]) for ncu captures; prints launches per step to stderr.
Never a bench value: numbers printed under a profiler are not timings."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import kmercamel_b200 as kb
from kmercamel_b200 import synth

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 1
recs = synth.random_genome_records(50, 1_000_000, 12345)
seq, off, ln = synth.frame_records(recs)
ctx = kb.Context(0, torch.cuda.current_stream().cuda_stream)
d = torch.from_numpy(seq).cuda()
for i in range(warm + steps):
    r = ctx.compute_device(d.data_ptr(), d.numel(), k=31)
    print(f"step {i}: launches={r.n_launches} kmers={r.n_kmers} len={r.length} times={r.times_ms}", file=sys.stderr)

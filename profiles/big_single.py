import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import kmercamel_b200 as kb
from kmercamel_b200 import synth
parts = []
for r in range(8):
    recs = synth.random_genome_records(50, 1_000_000, 12345 + r)
    parts.append(synth.frame_records(recs)[0])
seq = np.concatenate(parts)
ctx = kb.Context(0, torch.cuda.current_stream().cuda_stream)
d = torch.from_numpy(seq).cuda()
ctx.profile_enable(True)
for i in range(3):
    ctx.profile_reset()
    r = ctx.compute_device(d.data_ptr(), d.numel(), k=31)
    print(i, r.n_kmers, r.length, r.n_nodes, r.times_ms, file=sys.stderr)
print({k: round(v["ms"], 3) for k, v in ctx.profile().items() if v["launches"]}, file=sys.stderr)

import sys, time, numpy as np, torch
sys.path.insert(0,'.')
import kmercamel_b200 as kb
from kmercamel_b200 import synth
recs = synth.random_genome_records(50, 1_000_000, 12345)
seq, off, ln = synth.frame_records(recs)
ctx = kb.Context(0, torch.cuda.current_stream().cuda_stream)
d = torch.from_numpy(seq).cuda()
for i in range(4):
    t=time.perf_counter(); r = ctx.compute_device(d.data_ptr(), d.numel(), k=31); print('dev', (time.perf_counter()-t)*1000, r.times_ms['total'], file=sys.stderr)

import sys, time, numpy as np, torch
sys.path.insert(0,'.')
import kmercamel_b200 as kb
from kmercamel_b200 import synth
recs = synth.random_genome_records(50, 1_000_000, 12345)
seq, off, ln = synth.frame_records(recs)
ctx = kb.Context(0, torch.cuda.current_stream().cuda_stream)
pinned = torch.from_numpy(seq).pin_memory(); pn = pinned.numpy()
d = torch.from_numpy(seq).cuda()
for i in range(3):
    t=time.perf_counter(); r = ctx.compute_device(d.data_ptr(), d.numel(), k=31); print('dev', time.perf_counter()-t, r.times_ms['total'])
ctx.profile_enable(True); ctx.profile_reset()
for i in range(3):
    t=time.perf_counter(); r = ctx.compute_device(d.data_ptr(), d.numel(), k=31); print('dev prof', time.perf_counter()-t, r.times_ms['total'])
t=time.perf_counter(); p = ctx.profile(); print('profile()', time.perf_counter()-t)
ctx.profile_enable(False)
for i in range(5):
    t=time.perf_counter(); r = ctx.compute(pn, k=31, copy=False); print('host pinned', time.perf_counter()-t, r.times_ms['total'])

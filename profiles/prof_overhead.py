import os, sys, time
sys.path.insert(0, "/root/repo")
import torch
import kmercamel_b200 as kb
from kmercamel_b200 import synth
recs = synth.random_genome_records(50, 1_000_000, 12345)
seq, off, ln = synth.frame_records(recs)
st = torch.cuda.current_stream()
ctx = kb.Context(0, st.cuda_stream)
ds = [torch.from_numpy(seq).cuda() for _ in range(4)]
def run(n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(st)
    for i in range(n):
        r = ctx.compute_device(ds[i % 4].data_ptr(), ds[0].numel(), k=31)
    e1.record(st)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for i in range(5): ctx.compute_device(ds[i % 4].data_ptr(), ds[0].numel(), k=31)
for rep in range(3):
    ctx.profile_enable(False)
    a = run(30)
    ctx.profile_enable(True); ctx.profile_reset()
    b = run(30)
    print("no timers %.4f ms   with kernel timers %.4f ms" % (a, b))

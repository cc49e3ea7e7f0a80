"""Phase clocks of the single-CTA engine kernel (KC_TRACE) for genomes of 50 / 100 / 200 / 400 / 800 records, 256 vs 512 threads.
usage: KC_TRACE=1 python profiles/small_engine_trace.py 2> trace.log"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import kmercamel_b200 as kb
from kmercamel_b200 import synth

ctx = kb.Context(0, torch.cuda.current_stream().cuda_stream)
for n_rec in (50, 100, 200, 400, 800):
    seq, _, _ = synth.frame_records(synth.random_genome_records(n_rec, 40_000, 500 + n_rec))
    d = torch.from_numpy(seq).cuda()
    for thr in (256, 512):
        ctx.set_option("small_threads", thr)
        ctx.compute_device(d.data_ptr(), d.numel(), k=31)
        ctx.profile_enable(True); ctx.profile_reset()
        print(f"### records={n_rec} threads={thr}", file=sys.stderr, flush=True)
        for _ in range(3):
            r = ctx.compute_device(d.data_ptr(), d.numel(), k=31)
        p = ctx.profile()
        ctx.profile_enable(False)
        print(f"records={n_rec} threads={thr} small_engine_ms={p['small_engine']['ms'] / 3:.4f} path_ms={r.times_ms['path']:.4f}", flush=True)

"""Two passes of one of synth.BIG_CONFIGS through kc_compute_device, for ncu captures of the wide-word and 3.1 Gbp runs.
Never a bench value.  usage: step_for_ncu_cfg.py NAME"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import kmercamel_b200 as kb
from kmercamel_b200 import synth

name = sys.argv[1]
cfg = synth.BIG_CONFIGS[name]
seq, _, _ = synth.big_config_input(name)
ctx = kb.Context(0, torch.cuda.current_stream().cuda_stream)
d = torch.from_numpy(seq).cuda()
for i in range(2):
    r = ctx.compute_device(d.data_ptr(), d.numel(), k=cfg["k"], complements=cfg["complements"], min_frequency=cfg["min_frequency"])
    print(f"{name} pass {i}: launches={r.n_launches} kmers={r.n_kmers} len={r.length} times={r.times_ms} sig_runs={ctx.stat('sig_runs')}", file=sys.stderr)

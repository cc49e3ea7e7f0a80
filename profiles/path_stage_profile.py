"""Where the overlap stage spends its time on a many-node input: per-kernel-class device time against the stage time of
kc_compute_device, on the 310 Mbp scale model of configs[4] (126 k simplitigs, 31 levels).  Evidence for DESIGN.md §5: the stage is
bound by launches and host round trips per level, not by the tuple sort, which is why it is repeated on every rank instead of sharded.
usage: python profiles/path_stage_profile.py [config-name]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import kmercamel_b200 as kb
from kmercamel_b200 import synth

name = sys.argv[1] if len(sys.argv) > 1 else "cfg4_human_310M"
cfg = synth.BIG_CONFIGS[name]
seq, _, _ = synth.big_config_input(name)
d = torch.from_numpy(seq).cuda()
ctx = kb.Context(0, torch.cuda.current_stream().cuda_stream)
kw = dict(k=cfg["k"], complements=cfg["complements"], min_frequency=cfg["min_frequency"])
for _ in range(2):
    r = ctx.compute_device(d.data_ptr(), d.numel(), **kw)
ctx.profile_enable(True)
ctx.profile_reset()
steps = 3
for _ in range(steps):
    r = ctx.compute_device(d.data_ptr(), d.numel(), **kw)
prof = ctx.profile()
path_classes = ["tuples", "sort_hist", "sort_scatter", "sort_local", "sort_misc", "compact", "scan", "simulate", "doubling", "commit", "small_engine", "misc"]
out = {"config": cfg["config"], "n_bytes": int(d.numel()), "n_kmers": r.n_kmers, "nodes": r.n_nodes, "launches_per_step": r.n_launches,
       "stage_ms": r.times_ms,
       "kernel_classes_ms_per_step": {k: round(v["ms"] / steps, 4) for k, v in prof.items() if v["launches"]},
       "kernel_launches_per_step": {k: v["launches"] / steps for k, v in prof.items() if v["launches"]}}
out["path_kernels_ms"] = round(sum(prof[c]["ms"] for c in path_classes if c in prof) / steps, 3)
out["note"] = "path_kernels_ms also contains the sort_* / compact / scan / misc launches of the k-mer set stage when it took the exact path"
print(json.dumps(out))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "path_stage_%s.json" % name), "w"), indent=1)

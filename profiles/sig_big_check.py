"""Signature-bucket construction on the full-size configs (one GPU): which construction ran, stage times with it and with the
fixed-slot / exact constructions, identical results.  usage: python profiles/sig_big_check.py name [name ...]"""
import json, os, sys, hashlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import kmercamel_b200 as kb
from kmercamel_b200 import synth

out = []
ctx = kb.Context(0, torch.cuda.current_stream().cuda_stream)
for name in sys.argv[1:]:
    cfg = synth.BIG_CONFIGS[name]
    seq, _, _ = synth.big_config_input(name)
    d = torch.from_numpy(seq).cuda()
    kw = dict(k=cfg["k"], complements=cfg["complements"], min_frequency=cfg["min_frequency"])
    row = {"config": cfg["config"], "n_bytes": int(d.numel())}
    digests = []
    for label, sig in (("signature_buckets", 1), ("fixed_slots_or_exact", 0)):
        ctx.set_option("sig_set", sig)
        ctx.set_option("fast_heuristics", 1)
        r0 = (ctx.stat("sig_runs"), ctx.stat("sig_fallbacks"), ctx.stat("fast_runs"), ctx.stat("fast_fallbacks"))
        for _ in range(2):
            r = ctx.compute_device(d.data_ptr(), d.numel(), **kw)
        ctx.profile_enable(True)
        ctx.profile_reset()
        steps = 3
        for _ in range(steps):
            r = ctx.compute_device(d.data_ptr(), d.numel(), **kw)
        prof = ctx.profile()
        ctx.profile_enable(False)
        r1 = (ctx.stat("sig_runs"), ctx.stat("sig_fallbacks"), ctx.stat("fast_runs"), ctx.stat("fast_fallbacks"))
        ms = ctx.copy_to_host(r.ms_ptr, r.length)
        digests.append(hashlib.md5(ms).hexdigest())
        row[label] = {"stage_ms": {k: round(v, 3) for k, v in r.times_ms.items()}, "n_kmers": r.n_kmers, "length": r.length, "nodes": r.n_nodes,
                      "sig_runs/sig_fallbacks/fast_runs/fast_fallbacks over 5 calls": [b - a for a, b in zip(r0, r1)],
                      "set_kernels_ms": {k: round(v["ms"] / steps, 3) for k, v in prof.items() if k in ("ks_scatter0", "ks_resolve", "ks_hist0", "sort_scatter", "sort_hist", "sort_local") and v["launches"]}}
        del ms
    row["identical_superstrings"] = digests[0] == digests[1]
    ctx.set_option("sig_set", 1)
    out.append(row)
    print(json.dumps(row), flush=True)
    del d, seq
    torch.cuda.empty_cache()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "sig_big_check.json"), "w"), indent=1)

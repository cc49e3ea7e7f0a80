#!/usr/bin/env python3
"""One-call sweep of the fixed-slot construction's kernel variants (kc_set_option fast_resolve / fast_tile_variant).

For every (fast_resolve, fast_tile_variant) pair:
  * parity: the superstring of configs[1] and of a battery of smaller inputs (duplicate-heavy, -z 2 on the fast path,
    k = 63, k = 127 -u) must be byte-identical to the baseline variant (1, 0), which the GPU tests pin against the oracle;
    the fixed-slot path must really have run (fast_runs / fast_fallbacks counters);
  * time: device ms per step on configs[1] (CUDA events around 20 steps) + the per-kernel-class timers.
Writes gpurun_out/variant_sweep.json and prints the best passing pair as `KC_FAST_RESOLVE=r KC_FAST_TILE=t`.

    python profiles/variant_sweep.py [--steps 20]
"""
import argparse
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def battery():
    """(name, seq, kwargs, options) small parity inputs."""
    from kmercamel_b200 import synth
    out = []
    recs = synth.random_genome_records(4, 1_000_000, 777)
    comp = {65: 84, 67: 71, 71: 67, 84: 65}
    lut = np.zeros(256, dtype=np.uint8)
    for a, b in comp.items():
        lut[a] = b
    dup = recs + [recs[1].copy(), lut[recs[2]][::-1].copy(), recs[1][:500_000].copy()]  # every k-mer of r1 three times, r2 twice (as RC)
    seq, _, _ = synth.frame_records(dup)
    out.append(("dups_k31", seq, dict(k=31), {}))
    out.append(("dups_k31_u", seq, dict(k=31, complements=False), {}))
    out.append(("dups_k31_z2_fast", seq, dict(k=31, min_frequency=2), {"fast_heuristics": 0}))
    out.append(("dups_k31_z3_fast", seq, dict(k=31, min_frequency=3), {"fast_heuristics": 0}))
    out.append(("dups_k63", seq, dict(k=63), {}))
    out.append(("dups_k127_u", seq, dict(k=127, complements=False), {}))
    withn = seq.copy()
    withn[np.random.default_rng(5).integers(0, withn.size, size=2000)] = ord("N")
    out.append(("dups_withN_k23", withn, dict(k=23), {}))
    small, _, _ = synth.frame_records(synth.random_genome_records(3, 70_001, 99))
    out.append(("small_k15", small, dict(k=15), {}))
    return out


def run_variant(r, t, steps, ref, w=0):  # w = fast_max_ctas
    """One (fast_resolve, fast_tile_variant) pair in THIS process -> result row (ref: signatures of the baseline or {})."""
    import torch
    import kmercamel_b200 as kb
    import bench

    _, (seq, _, _) = bench.make_workload(0)
    stream = torch.cuda.current_stream()
    ctx = kb.Context(0, stream.cuda_stream)
    d_seq = torch.from_numpy(seq).cuda()
    row = {"fast_resolve": r, "fast_tile_variant": t, "fast_max_ctas": w, "ok": True, "cases": {}}
    if r < 0:  # the exact (histogram-based) construction of kmerset.cuh: an independent implementation as the reference
        ctx.set_option("fast_set", 0)
    else:
        ctx.set_option("fast_resolve", r)
        ctx.set_option("fast_tile_variant", t)
        ctx.set_option("fast_max_ctas", w)
    cases = [("configs1", seq, dict(k=bench.K), {})] + battery()
    for name, s, kw, opts in cases:
        for o, v in opts.items():
            ctx.set_option(o, v)
        f0, b0 = ctx.stat("fast_runs"), ctx.stat("fast_fallbacks")
        rr = ctx.compute(s, **kw)
        sg = [hashlib.md5(rr.ms).hexdigest(), rr.length, rr.n_kmers]
        row["cases"][name] = {"sig": sg, "fast": ctx.stat("fast_runs") - f0, "fallbacks": ctx.stat("fast_fallbacks") - b0}
        if ref and sg != ref[name]["sig"]:
            row["ok"] = False
            row.setdefault("mismatch", []).append(name)
        ctx.set_option("fast_heuristics", 1)
    for _ in range(3):
        ctx.compute_device(d_seq.data_ptr(), d_seq.numel(), k=bench.K)
    ctx.profile_enable(True)
    ctx.profile_reset()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        ctx.compute_device(d_seq.data_ptr(), d_seq.numel(), k=bench.K)
    e1.record(stream)
    torch.cuda.synchronize()
    row["ms_per_step"] = e0.elapsed_time(e1) / steps
    prof = ctx.profile()
    ctx.profile_enable(False)
    row["kernel_ms"] = {n: round(v["ms"] / steps, 4) for n, v in prof.items() if v["launches"]}
    return row


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "variant_sweep.json"))
    ap.add_argument("--variant", default=None, help="r,t: run this pair here and print its row (used by the parent)")
    ap.add_argument("--ref", default=None)
    args = ap.parse_args()
    if args.variant:
        r, t, w = (int(x) for x in args.variant.split(","))
        ref = json.load(open(args.ref)) if args.ref and os.path.exists(args.ref) else {}
        print("ROW " + json.dumps(run_variant(r, t, args.steps, ref, w)), flush=True)
        return
    # parent: one process per variant (a faulting kernel poisons its CUDA context), each under a timeout
    import subprocess
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    ref_path = args.out + ".ref"
    if os.path.exists(ref_path):
        os.remove(ref_path)
    # (fast_resolve, fast_tile_variant, fast_max_ctas); the first one supplies the reference signatures (-1 = exact construction)
    variants = [(-1, 0, 0), (1, 0, 0), (6, 5, 2368)]
    rows = []
    for (r, t, w) in variants:
        row = {"fast_resolve": r, "fast_tile_variant": t, "fast_max_ctas": w, "ok": False}
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), f"--variant={r},{t},{w}", "--steps", str(args.steps), "--ref", ref_path],
                               capture_output=True, text=True, timeout=150)
            got = [ln for ln in p.stdout.splitlines() if ln.startswith("ROW ")]
            if got:
                row = json.loads(got[-1][4:])
            else:
                row["error"] = (p.stderr or "")[-600:]
        except subprocess.TimeoutExpired:
            row["error"] = "timeout"
        if (r, t, w) == variants[0]:
            if not row.get("ok"):
                print("baseline variant failed: " + str(row.get("error")), file=sys.stderr)
            else:
                json.dump(row["cases"], open(ref_path, "w"))
        rows.append(row)
        print(json.dumps({k: v for k, v in row.items() if k != "cases"}), file=sys.stderr, flush=True)
        with open(args.out, "w") as f:
            json.dump({"rows": rows}, f, indent=1)
    good = [x for x in rows if x.get("ok") and "ms_per_step" in x]
    best = min(good, key=lambda x: x["ms_per_step"]) if good else None
    with open(args.out, "w") as f:
        json.dump({"rows": rows, "best": best and [best["fast_resolve"], best["fast_tile_variant"], best["fast_max_ctas"]]}, f, indent=1)
    if best:
        print(f"KC_FAST_RESOLVE={best['fast_resolve']} KC_FAST_TILE={best['fast_tile_variant']} KC_FAST_MAX_CTAS={best['fast_max_ctas']}")
    else:
        print("KC_FAST_RESOLVE=1 KC_FAST_TILE=0 KC_FAST_MAX_CTAS=0")


if __name__ == "__main__":
    main()

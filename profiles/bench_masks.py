"""Throughput of the compute path's neighbours on the BASELINE configs[1] input (50 x 1 Mbp,
k = 31): kc_streaming and kc_maskopt through the C ABI with HOST buffers (H2D + kernels + D2H inside the timed region),
next to the unmodified reference CLI (oracle/_ref/kmercamel, one host core) on a bounded sample.  One JSON object on stdout.
Diagnosis / reporting only: the driver's bench contract is bench.py."""
import json, os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import kmercamel_b200 as kb
from kmercamel_b200 import synth

K = 31
REF = os.path.join(ROOT, "oracle", "_ref", "kmercamel")


def timed(fn, steps=5, warm=2):
    for _ in range(warm):
        r = fn()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        r = fn()
        ts.append(time.perf_counter() - t0)
    return r, float(np.median(ts)) * 1e3


def main():
    import torch
    recs = synth.random_genome_records(50, 1_000_000, 12345)
    seq0, off, ln = synth.frame_records(recs)
    pin = torch.from_numpy(seq0).pin_memory()   # host buffers are pinned, results stay in the library's pinned buffer (copy=False)
    seq = pin.numpy()
    ctx = kb.Context(0)
    out = {"workload": "BASELINE configs[1]: 50 x 1 Mbp uniform ACGT, k=31 canonical", "k": K, "bases": int(seq.size)}

    ctx.profile_enable(True)
    r, ms = timed(lambda: ctx.streaming(seq, k=K, copy=False))
    prof = {k: v for k, v in ctx.profile().items() if v["launches"]}
    ctx.profile_enable(False)
    out["streaming"] = {"e2e_ms": ms, "device_ms": r.times_ms["total"], "length": r.length, "kmers_on": r.n_kmers, "launches": r.n_launches,
                        "kmers_per_s_e2e": r.n_kmers / (ms * 1e-3), "kmers_per_s_device": r.n_kmers / (r.times_ms["total"] * 1e-3)}
    r2, ms2 = timed(lambda: ctx.streaming(seq, k=K, min_frequency=2, copy=False), steps=3, warm=1)
    out["streaming_z2"] = {"e2e_ms": ms2, "device_ms": r2.times_ms["total"], "length": r2.length, "launches": r2.n_launches}

    # maskopt input: the greedy superstring of the same genome with its default (min-one) mask
    comp = ctx.compute(seq, k=K)
    ms_pin = torch.frombuffer(bytearray(comp.ms), dtype=torch.uint8).pin_memory()
    ms_in = ms_pin.numpy()
    for name, minimize in (("maskopt_maxone", False), ("maskopt_minone", True)):
        m, t = timed(lambda: ctx.maskopt(ms_in, k=K, minimize=minimize, copy=False), steps=3, warm=1)
        ones = int(np.count_nonzero(np.frombuffer(ctx.maskopt(ms_in, k=K, minimize=minimize).ms, dtype=np.uint8) <= 90))
        out[name] = {"e2e_ms": t, "device_ms": m.times_ms["total"], "length": m.length, "set_size": m.n_kmers, "ones": ones,
                     "launches": m.n_launches, "chars_per_s_e2e": m.length / (t * 1e-3)}
    assert out["maskopt_minone"]["ones"] == out["maskopt_minone"]["set_size"] == comp.n_kmers

    if os.path.exists(REF):  # reference CLI, one core, first 5 records (5 Mbp)
        with tempfile.TemporaryDirectory() as td:
            fa = os.path.join(td, "s.fa")
            open(fa, "wb").write(synth.fasta_bytes(recs[:5]))
            t0 = time.perf_counter()
            p = subprocess.run([REF, "compute", "-a", "streaming", "-k", str(K), fa], capture_output=True)
            t_stream = time.perf_counter() - t0
            msf = os.path.join(td, "ms.fa")
            open(msf, "wb").write(p.stdout)
            n_on = sum(1 for c in p.stdout.split(b"\n")[1] if c <= 90)
            cpu = {"sample": "first 5 of the 50 records (5 Mbp), one host core", "streaming_s": t_stream, "streaming_kmers_per_s": n_on / t_stream}
            for t in ("max-one", "min-one"):
                t0 = time.perf_counter()
                subprocess.run([REF, "maskopt", "-t", t, "-k", str(K), msf], capture_output=True)
                dt = time.perf_counter() - t0
                cpu["maskopt_%s_s" % t] = dt
                cpu["maskopt_%s_chars_per_s" % t] = len(p.stdout.split(b"\n")[1]) / dt
            out["cpu_reference"] = cpu
    out["streaming_kernel_classes"] = {k: {"ms_per_call": v["ms"] / 7, "launches_per_call": v["launches"] / 7,
                                           "gbs": (v["bytes"] / 7) / (v["ms"] / 7 * 1e-3) / 1e9 if v["ms"] and v["bytes"] else None} for k, v in prof.items()}
    print(json.dumps(out))


if __name__ == "__main__":
    main()

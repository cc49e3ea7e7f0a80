"""Wall clock of the CLI (file read + framing + CUDA context + H2D + kernels + D2H + file write) next to the unmodified
reference CLI on the same FASTA file (BASELINE configs[1], 50 x 1 Mbp, k = 31), both on this box.  One JSON object on stdout."""
import json, os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kmercamel_b200 import synth

OURS = os.path.join(ROOT, "host", "kmercamel")
REF = os.path.join(ROOT, "oracle", "_ref", "kmercamel")


def wall(cmd):
    t0 = time.perf_counter()
    p = subprocess.run(cmd, capture_output=True)
    return time.perf_counter() - t0, p


with tempfile.TemporaryDirectory() as td:
    fa = os.path.join(td, "cfg2.fa")
    open(fa, "wb").write(synth.fasta_bytes(synth.random_genome_records(50, 1_000_000, 12345)))
    out = {"workload": "BASELINE configs[1]: 50 x 1 Mbp multi-FASTA (80-column lines), k=31 canonical", "file_bytes": os.path.getsize(fa), "nproc": os.cpu_count()}
    runs = []
    for i in range(4):
        o = os.path.join(td, "o%d.fa" % i)
        t, p = wall([OURS, "compute", "-k", "31", "-o", o, fa])
        assert p.returncode == 0, p.stderr[-400:]
        runs.append(t)
        stages = [l for l in p.stderr.decode().splitlines() if "GPU stages" in l]
    line = open(o, "rb").read().split(b"\n")[1]
    out["ours"] = {"wall_s": runs, "best_wall_s": min(runs), "length": len(line), "ones": sum(1 for c in line if c <= 90), "gpu_stages": stages[-1].split("] ", 1)[1]}
    tm, p = wall([OURS, "compute", "-k", "31", "-M", os.path.join(td, "mo.fa"), "-o", os.path.join(td, "om.fa"), fa])
    out["ours_with_maxone"] = {"wall_s": tm}
    ts, p = wall([OURS, "compute", "-a", "streaming", "-k", "31", "-o", os.path.join(td, "os.fa"), fa])
    out["ours_streaming"] = {"wall_s": ts}
    if os.path.exists(REF):
        t, p = wall([REF, "compute", "-k", "31", "-o", os.path.join(td, "r.fa"), fa])
        rl = open(os.path.join(td, "r.fa"), "rb").read().split(b"\n")[1]
        out["reference"] = {"wall_s": t, "cores": 1, "length": len(rl), "ones": sum(1 for c in rl if c <= 90)}
        out["wall_clock_ratio"] = t / min(runs)
        t, p = wall([REF, "compute", "-a", "streaming", "-k", "31", "-o", os.path.join(td, "rs.fa"), fa])
        out["reference_streaming"] = {"wall_s": t}
        out["streaming_identical"] = open(os.path.join(td, "rs.fa"), "rb").read().split(b"\n")[1] == open(os.path.join(td, "os.fa"), "rb").read().split(b"\n")[1]
    print(json.dumps(out))

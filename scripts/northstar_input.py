"""Writes BASELINE configs[4] (3.1 Gbp repeat-model genome, synth.human_like_genome(3_100_000_000, 3100)) — or its 310 Mbp
scale model with `small` — as a FASTA file, one line per record (line structure does not change the k-mers; the reference's
golden run in tests/golden/make_golden_big.py used 80-column lines of the same records).  usage: northstar_input.py OUT [small]"""
import os
import sys
import time
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kmercamel_b200 import synth  # noqa: E402

t0 = time.time()
n = 310_000_000 if len(sys.argv) > 2 and sys.argv[2] == "small" else 3_100_000_000
recs = synth.human_like_genome(n, 3100)
crc = 0
with open(sys.argv[1], "wb") as f:
    for i, r in enumerate(recs):
        f.write(b">r%d_0\n" % i)
        r.tofile(f)
        f.write(b"\n")
        mv = memoryview(r)
        for lo in range(0, len(r), 1 << 28):
            crc = zlib.crc32(mv[lo:lo + (1 << 28)], crc)
        crc = zlib.crc32(b"\n", crc)
print("northstar input: %d bases in %d records, crc32 of the framed sequence %d, %.1f s" % (n, len(recs), crc, time.time() - t0))

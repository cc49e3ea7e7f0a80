import sys, numpy as np
sys.path.insert(0,'.')
import kmercamel_b200 as kb
from oracle import orc
ctx = kb.Context(0)
import gzip
seq,_,_ = kb.frame_fasta(gzip.open('tests/golden/spneumoniae.fa.gz').read())
keys,vals = ctx.count_kmers(seq,k=31)
print(len(keys))
keys,vals = ctx.count_kmers(seq,k=127,complements=False)
print(len(keys))
rng = np.random.default_rng(1)
for k, n, compl in [(9, 3000, True), (15, 5000, False)]:
    L = orc.limbs_for_k(k)
    genome = rng.integers(0, 4, size=20000)
    def word(pos):
        v = 0
        for c in genome[pos:pos + k]:
            v = (v << 2) | int(c)
        return [(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(L)]
    starts = rng.integers(0, 20000 - k - 40, size=n)
    lens = rng.integers(0, 30, size=n)
    first = np.array([word(int(s)) for s in starts], dtype=np.uint64)
    last = np.array([word(int(s + l)) for s, l in zip(starts, lens)], dtype=np.uint64)
    want_ef, want_ov = orc.overlap_path(first, last, k, compl)
    ef, ov = ctx.overlap_path(first, last, k=k, complements=compl, strict=True)
    print(k, n, compl, np.array_equal(ef, want_ef), np.array_equal(ov, want_ov))

"""GPU: the FULL-SIZE BASELINE.json configurations against the reference's own results (tests/golden/golden_big.json, produced
in the build container by tests/golden/make_golden_big.py from the unmodified reference, oracle/_ref/ref_harness `full`).

Per configuration the input is regenerated from its seed (crc32 pinned), then:
  * kc_kmer_digest(input)            == the reference's digest of its hash table (n, sum h, xor h, sum h * count): the k-mer set
                                        and the -z counts are bit-exact without moving 10^9 keys;
  * kc_compute(input).n_kmers        == the reference's count; ones == n_kmers;
  * kc_kmer_digest(masked, output)   == the reference's (n, sum h, xor h): the superstring represents exactly the reference's
                                        k-mer set — the check of the reference's verify.py;
  * length within 0.1 % of the reference's superstring (north star; the reference's simplitig order is unspecified).
KC_SKIP_BIG=1 skips the configurations above 1 Gbase."""
import json
import os
import zlib

import numpy as np
import pytest

from kmercamel_b200 import synth

from conftest import GOLDEN_DIR

pytestmark = pytest.mark.gpu
BIG = json.load(open(os.path.join(GOLDEN_DIR, "golden_big.json")))
HUGE = {"cfg3_reads", "cfg4_human"}
ORDER = ["tiny_k31z2", "tiny_k63", "tiny_k127u", "cfg1_50M", "cfg3_reads_10M", "cfg4_human_310M", "cfg2_k63u", "cfg2_k127u", "cfg3_reads",
         "cfg4_human"]


def _crc(seq):
    c, mv = 0, memoryview(seq)
    for lo in range(0, len(seq), 1 << 28):
        c = zlib.crc32(mv[lo:lo + (1 << 28)], c)
    return c


@pytest.mark.parametrize("name", [n for n in ORDER if n in BIG])
def test_big_config_against_reference(ctx, name):
    if name in HUGE and os.environ.get("KC_SKIP_BIG"):
        pytest.skip("KC_SKIP_BIG")
    g = BIG[name]
    ref = g["reference"]
    k, compl, z = g["k"], g["complements"], g["min_frequency"]
    seq, _, _ = synth.big_config_input(name)
    assert len(seq) == g["n_bytes"] and _crc(seq) == g["crc32"], "the generator no longer reproduces the golden input"
    d = ctx.kmer_digest(seq, k=k, complements=compl, min_frequency=z)
    assert d == ref["digest"], "k-mer set / counts differ from the reference"
    r = ctx.compute(seq, k=k, complements=compl, min_frequency=z, copy=False)
    del seq
    assert r.n_kmers == ref["n_kmers"]
    if "length" in ref:
        assert abs(r.length - ref["length"]) <= 1e-3 * ref["length"], (r.length, ref["length"])
    import ctypes as C
    ms = np.ctypeslib.as_array(C.cast(r.ms_ptr, C.POINTER(C.c_uint8)), shape=(r.length,))  # the context's pinned result buffer
    assert int((ms <= 90).sum()) == ref["n_kmers"] and bool((ms[r.length - (k - 1):] > 90).all())
    ms = ms.copy()                                                                         # the next call reuses that buffer
    dm = ctx.kmer_digest(ms, k=k, complements=compl, masked=True)
    assert dm[:3] == ref["digest"][:3], "the superstring does not represent the reference's k-mer set"

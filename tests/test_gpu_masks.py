"""GPU parity tests of the compute path's neighbours, through the C ABI:
kc_streaming (`compute -a streaming [-z]`, reference src/streaming.h:12-107) and kc_maskopt (`maskopt -t max-one|min-one`,
reference src/masks.h:40-78,240-261) against outputs of the unmodified reference CLI (tests/golden/golden_masks.json)
and, on seeded random inputs, against the oracle restatement."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR, md5
from kmercamel_b200 import api, synth
from oracle import orc
from test_oracle_masks import framed, parse_flags, reads_fasta

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gm():
    return json.load(open(os.path.join(GOLDEN_DIR, "golden_masks.json")))


def test_streaming_unittest_vectors(ctx, test_fa_bytes):
    seq, _, _ = framed(test_fa_bytes)
    for k, c, want in [(3, True, "ACCCGAacCGtaATgcTTta"), (2, True, "ACcCGAaTaATGc"), (1, True, "AC"), (1, False, "ACGT")]:
        assert ctx.streaming(seq, k=k, complements=c).ms.decode() == want  # tests/streaming_unittest.h:18-23
    for k, c, z, want in [(3, True, 2, "ACCCGttTaa"), (3, False, 2, "ACCCgtAac"), (3, True, 3, "AAcg"), (4, True, 2, "ACccgAacg"),
                          (1, True, 2, "CA"), (1, False, 2, "CAGT"), (1, False, 3, "CAGT"), (1, False, 4, "CAGT"), (1, False, 5, "CATG"),
                          (1, False, 6, "CA"), (1, False, 10, "C")]:
        assert ctx.streaming(seq, k=k, complements=c, min_frequency=z).ms.decode() == want  # tests/streaming_unittest.h:44-56


def test_streaming_fuzz_vs_reference(ctx, gm):
    for case in gm["streaming_fuzz"]:
        fasta = "".join(">r%d\n%s\n" % (j, r) for j, r in enumerate(case["records"])).encode()
        seq, _, _ = framed(fasta)
        r = ctx.streaming(seq, k=case["k"], complements=not case["unidirectional"], min_frequency=case["z"])
        assert r.ms.decode() == case["seq"], case
        assert r.n_kmers == sum(1 for ch in case["seq"] if ch <= "Z")


def test_streaming_test_fa_vs_reference(ctx, gm, test_fa_bytes):
    seq, _, _ = framed(test_fa_bytes)
    for name, row in gm["streaming"].items():
        if name.startswith("test_"):
            k, c, z = parse_flags(name)
            assert ctx.streaming(seq, k=k, complements=c, min_frequency=z).ms.decode() == row["seq"], name


@pytest.mark.parametrize("name", ["sp_k31", "sp_k31u", "sp_k13", "sp_k63", "sp_k127u", "sp_k31z2", "sp_k5z3"])
def test_streaming_spneumoniae_vs_reference(ctx, gm, spneumoniae_bytes, name):
    seq, _, _ = framed(spneumoniae_bytes)
    k, c, z = parse_flags(name)
    r = ctx.streaming(seq, k=k, complements=c, min_frequency=z)
    row = gm["streaming"][name]
    assert (r.length, md5(r.ms)) == (row["length"], row["md5"])
    assert r.n_kmers == row["ones"]


@pytest.mark.parametrize("name", ["reads_k31", "reads_k31z2", "reads_k21z3u", "reads_k47z2"])
def test_streaming_reads_vs_reference(ctx, gm, name):
    seq, _, _ = framed(reads_fasta(gm["reads_params"]))
    k, c, z = parse_flags(name)
    r = ctx.streaming(seq, k=k, complements=c, min_frequency=z)
    row = gm["streaming"][name]
    assert (r.length, md5(r.ms)) == (row["length"], row["md5"])


def test_streaming_simplitigs_vs_reference(ctx, gm, simplitigs_bytes):
    seq, _, _ = framed(simplitigs_bytes)
    r = ctx.streaming(seq, k=31)
    row = gm["streaming"]["sim_k31"]
    assert (r.length, md5(r.ms)) == (row["length"], row["md5"])


@pytest.mark.parametrize("k,c,z", [(31, True, 1), (15, False, 1), (63, True, 2), (100, False, 1), (9, True, 4)])
def test_streaming_random_reads_vs_oracle(ctx, k, c, z):
    """Larger seeded inputs (multi-level fast plan, exact fallback for the read sets) against the oracle."""
    reads = synth.reads_from_genome(300_000, 8.0, 150, 0.01, 1000 + k)
    recs = [r for r in reads]
    recs.append(np.frombuffer(b"ACGTNNNNACGTTTGACCA" * 50, dtype=np.uint8))
    seq, off, ln = synth.frame_records(recs)
    want = orc.streaming(seq, off, ln, k, c, z)
    r = ctx.streaming(seq, k=k, complements=c, min_frequency=z)
    assert r.length == len(want) and r.ms == want


def test_streaming_edge_inputs(ctx):
    assert ctx.streaming(np.frombuffer(b"\n", dtype=np.uint8), k=5).length == 0
    assert ctx.streaming(np.frombuffer(b"ACG\nNNNN\n", dtype=np.uint8), k=5).length == 0
    assert ctx.streaming(np.frombuffer(b"ACGTA\n", dtype=np.uint8), k=5).ms == b"Acgta"
    assert ctx.streaming(np.frombuffer(b"acgta\nACGTA\nTACGT\n", dtype=np.uint8), k=5).ms == b"Acgta"
    assert ctx.streaming(np.frombuffer(b"acgta\nACGTA\nTACGT\n", dtype=np.uint8), k=5, complements=False).ms == b"AcgtaTacgt"
    with pytest.raises(api.KcError):
        ctx.streaming(np.frombuffer(b"ACGT\n", dtype=np.uint8), k=0)


def test_maskopt_fuzz_vs_reference(ctx, gm):
    for case in gm["maskopt_fuzz"]:
        for t in ("max-one", "min-one"):
            r = ctx.maskopt(case["ms"].encode(), k=case["k"], complements=not case["unidirectional"], minimize=t == "min-one")
            assert r.ms.decode() == case[t], (case, t)


def test_maskopt_simplitigs_vs_reference(ctx, gm, simplitigs_bytes):
    seq, off, ln = framed(simplitigs_bytes)
    ms = api.spss_to_ms(seq, off, ln, 31)
    for name, k, c in [("k31", 31, True), ("k31u", 31, False), ("k25", 25, True), ("k40u", 40, False)]:
        for t in ("max-one", "min-one"):
            r = ctx.maskopt(ms, k=k, complements=c, minimize=t == "min-one")
            want = gm["maskopt"]["sim_%s_%s" % (name, t)]
            assert (sum(1 for ch in r.ms if ch <= 90), md5(r.ms)) == (want["ones"], want["md5"]), (name, t)


@pytest.mark.parametrize("k,c,tag", [(31, True, "sp_stream_k31"), (70, False, "sp_stream_k70u")])
def test_maskopt_of_streaming_superstring_vs_reference(ctx, gm, spneumoniae_bytes, k, c, tag):
    """streaming -> maskopt chained on the GPU: both steps must reproduce the reference byte for byte."""
    seq, _, _ = framed(spneumoniae_bytes)
    ms = ctx.streaming(seq, k=k, complements=c).ms
    for t in ("max-one", "min-one"):
        r = ctx.maskopt(ms, k=k, complements=c, minimize=t == "min-one")
        want = gm["maskopt"]["%s_%s" % (tag, t)]
        assert (sum(1 for ch in r.ms if ch <= 90), md5(r.ms)) == (want["ones"], want["md5"]), t


def test_maskopt_properties_and_errors(ctx):
    rng = np.random.default_rng(5)
    base = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=200_000)]
    base[100_000:150_000] = base[:50_000]                       # repeats: max-one has more ones than min-one
    lower = rng.random(base.size) < 0.5
    ms = np.where(lower, base | 0x20, base).astype(np.uint8).tobytes()
    for k, c in [(21, True), (33, False), (90, True)]:
        mx = ctx.maskopt(ms, k=k, complements=c, minimize=False)
        mn = ctx.maskopt(ms, k=k, complements=c, minimize=True)
        assert mx.ms == orc.maskopt(ms, k, c, False) and mn.ms == orc.maskopt(ms, k, c, True)
        kx, _ = orc.ms_kmers(mx.ms, k, c)
        kn, on = orc.ms_kmers(mn.ms, k, c)
        k0, _ = orc.ms_kmers(ms, k, c)
        assert np.array_equal(kx, kn) and np.array_equal(kx, k0)    # the represented set is unchanged
        assert on == len(kn) == mn.n_kmers                           # min-one: one ON position per k-mer
        assert ctx.maskopt(mx.ms, k=k, complements=c, minimize=False).ms == mx.ms  # idempotent
    with pytest.raises(api.KcError) as e:
        ctx.maskopt(b"ACGTNACGT", k=3)
    assert e.value.code == -5
    assert ctx.maskopt(b"ACgt", k=7).ms == b"acgt"      # shorter than k: only the trailing lower-case part
    assert ctx.maskopt(b"", k=3).length == 0

"""Parity tests proper: the CUDA path, called through the C ABI, against the oracle and the reference goldens.

Run with `pytest -m gpu` on the B200 box.  Nothing here reads /root/reference.
"""
import numpy as np
import pytest

from kmercamel_b200 import synth
import kmercamel_b200 as kb
from oracle import orc
from conftest import keys_md5, md5

pytestmark = pytest.mark.gpu


def K(s, k=None):
    return orc.kmer_from_string(s, k)


# ---- stage 1: canonical k-mer set and counts, bit-exact -------------------------------------------------------
@pytest.mark.parametrize("name", ["k3", "k3u", "k10", "k5", "k5u", "k2", "k2u", "k1u", "k4"])
def test_count_kmers_test_fa(ctx, golden, test_fa_bytes, name):
    g = golden["test_fa_kmers"][name]
    k = int(name[1:].rstrip("u"))
    seq, off, ln = kb.frame_fasta(test_fa_bytes)
    keys, vals = ctx.count_kmers(seq, k=k, complements=not name.endswith("u"))
    assert len(keys) == g["n"] and keys_md5(keys) == g["keys_md5"] and md5(vals.tobytes()) == g["vals_md5"]


@pytest.mark.parametrize("k,compl,z,size", [(3, True, 2, 6), (3, False, 2, 5), (3, True, 3, 2), (4, True, 2, 3), (1, False, 5, 4),
                                            (1, False, 6, 2), (1, False, 10, 1)])
def test_count_kmers_filtered_sizes(ctx, test_fa_bytes, k, compl, z, size):
    """reference tests/parser_unittest.h:64-91 (ReadKMersFiltered)"""
    seq, _, _ = kb.frame_fasta(test_fa_bytes)
    keys, vals = ctx.count_kmers(seq, k=k, complements=compl, min_frequency=z)
    assert len(keys) == size and all(int(v) + 1 >= z for v in vals)


@pytest.mark.parametrize("name", ["k31", "k31u", "k63", "k127u", "k32", "k64", "k5", "k1u"])
def test_count_kmers_spneumoniae(ctx, golden, spneumoniae_bytes, name):
    g = golden["spneumoniae_kmers"][name]
    k = int(name[1:].rstrip("u"))
    seq, _, _ = kb.frame_fasta(spneumoniae_bytes)
    keys, vals = ctx.count_kmers(seq, k=k, complements=not name.endswith("u"))
    assert len(keys) == g["n"] and keys.shape[1] == g["limbs"]
    assert keys_md5(keys) == g["keys_md5"] and md5(vals.tobytes()) == g["vals_md5"]


def test_count_kmers_edge_inputs(ctx):
    # no k-mer at all, a single k-mer, only separators, windows broken by N and by record boundaries
    for text, k in [(b"\n", 3), (b"AC\n", 3), (b"ACG\n", 3), (b"NNNN\n\n\n", 2), (b"ACGNACG\nAC\nGT\n", 3), (b"A" * 100 + b"\n", 31)]:
        seq = np.frombuffer(text, dtype=np.uint8)
        recs = [r for r in text.split(b"\n")]
        s2, off, ln = orc.records_to_arrays(recs)
        want_k, want_v = orc.count_kmers(s2, off, ln, k, True)
        keys, vals = ctx.count_kmers(seq, k=k, complements=True)
        assert np.array_equal(keys, want_k) and np.array_equal(vals, want_v), text


def test_count_kmers_heavy_duplicates_and_saturation(ctx):
    """Buckets that never split (one k-mer repeated) and the 255 saturation of src/parser.h:77."""
    text = (b"A" * 40000 + b"\n") + (b"ACGTTGCA" * 3000 + b"\n") + (b"T" * 300 + b"\n")
    seq = np.frombuffer(text, dtype=np.uint8)
    s2, off, ln = orc.records_to_arrays(text.split(b"\n")[:-1])
    for k, compl in [(5, True), (21, False), (31, True), (40, True)]:
        want_k, want_v = orc.count_kmers(s2, off, ln, k, compl)
        keys, vals = ctx.count_kmers(seq, k=k, complements=compl)
        assert np.array_equal(keys, want_k) and np.array_equal(vals, want_v)
        assert vals.max() == 255


@pytest.mark.parametrize("k,compl,z", [(15, True, 1), (31, True, 1), (31, False, 2), (47, True, 1), (95, False, 1), (127, True, 3)])
def test_count_kmers_random_reads_vs_oracle(ctx, k, compl, z):
    reads = synth.reads_from_genome(60000, 12.0, 150, 0.01, seed=k)
    seq, off, ln = synth.frame_records(list(reads))
    want_k, want_v = orc.count_kmers(seq, off, ln, k, compl)
    keep = want_v.astype(int) + 1 >= z
    keys, vals = ctx.count_kmers(seq, k=k, complements=compl, min_frequency=z)
    assert np.array_equal(keys, want_k[keep]) and np.array_equal(vals, want_v[keep])


# ---- overlap stage: the reference's tie order ----------------------------------------------------------------
def test_overlap_path_kats(ctx):
    """reference tests/global_unittest.h:61-104 incl. lower_bound"""
    from test_oracle import PATH_KATS, node_ends
    for records, k, compl, lb, want_ef, want_ov in PATH_KATS:
        first, last = node_ends(records, k)
        ef, ov = ctx.overlap_path(first, last, k=k, complements=compl, lower_bound=lb, strict=True)
        assert ef.tolist() == want_ef and ov.tolist() == want_ov, records


def test_overlap_path_sparse_kats(ctx):
    """reference tests/global_sparse_unittest.h:84-103 (k-mer nodes: first == last)"""
    cases = [
        (["AT"], 2, True, [-1, -1], [255, 255]),
        (["ACG", "TAC", "GGC"], 3, False, [2, 0, -1], [1, 2, 255]),
        (["ACAA", "ATTT", "AACA"], 4, True, [4, 3, 0, 5, -1, -1], [2, 2, 3, 3, 255, 255]),
    ]
    for kmers, k, compl, want_ef, want_ov in cases:
        f = np.stack([K(x, k) for x in kmers])
        ef, ov = ctx.overlap_path(f, f, k=k, complements=compl, strict=True)
        assert ef.tolist() == want_ef and ov.tolist() == want_ov


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_overlap_path_random_vs_oracle(ctx, seed):
    rng = np.random.default_rng(seed)
    for k, n, compl in [(9, 3000, True), (15, 5000, False), (31, 2000, True), (40, 1500, True), (80, 800, False)]:
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
        assert np.array_equal(ef, want_ef) and np.array_equal(ov, want_ov), (k, n, compl)


# ---- lowerbound: kc_lower_bound against the reference CLI and its unit tests ------------------------
def test_lower_bound_kats(ctx):
    """reference tests/lower_bound_unittest.h:11-41"""
    from test_oracle import LOWER_BOUND_KATS
    for records, k, compl, want in LOWER_BOUND_KATS:
        seq, off, ln = orc.records_to_arrays([r.encode() for r in records])
        lb, _ = ctx.lower_bound(seq, off, ln, k=k, complements=compl, assume_simplitigs=True)
        assert lb == want, records


def test_lower_bound_S_exact(ctx, golden, simplitigs_bytes):
    """-S inputs: the value is a pure function of the record order -> equal to `kmercamel lowerbound -S` of the reference"""
    for g, want in zip(golden["fuzz_S"], golden["lowerbound"]["fuzz_S"]):
        seq, off, ln = orc.records_to_arrays([r.encode() for r in g["records"]])
        lb, _ = ctx.lower_bound(seq, off, ln, k=g["k"], complements=g["complements"], assume_simplitigs=True)
        assert lb == want, g
    seq, off, ln = kb.frame_fasta(simplitigs_bytes)
    for name, want in golden["lowerbound"]["simplitigs_S"].items():
        k = int(name[1:].rstrip("u"))
        lb, st = ctx.lower_bound(seq, off, ln, k=k, complements=not name.endswith("u"), assume_simplitigs=True)
        assert lb == want and st.n_nodes == len(off), name


@pytest.mark.parametrize("name", ["k31", "k31u", "k13", "k63", "k31z2"])
def test_lower_bound_from_fasta(ctx, golden, spneumoniae_bytes, name):
    """From FASTA the reference's value depends on its khash-order simplitigs: within 0.1 %, and never above the
    length of the superstring computed from the same input (it is a lower bound of it)."""
    want = golden["lowerbound"]["spneumoniae"][name]
    flags = name.lstrip("k0123456789")
    k = int(name[1:len(name) - len(flags)])
    z = 2 if "z2" in flags else 1
    seq, _, _ = kb.frame_fasta(spneumoniae_bytes)
    lb, st = ctx.lower_bound(seq, k=k, complements="u" not in flags, min_frequency=z)
    assert abs(lb - want) <= 0.001 * want, (lb, want)
    r = ctx.compute(seq, k=k, complements="u" not in flags, min_frequency=z)
    assert st.n_kmers == r.n_kmers and lb <= r.length


# ---- -S regime: byte-exact superstring and max-one mask ---------------------------------------------------------
def test_compute_S_fuzz_byte_exact(ctx, golden):
    for g in golden["fuzz_S"]:
        seq, off, ln = orc.records_to_arrays([r.encode() for r in g["records"]])
        r = ctx.compute(seq, off, ln, k=g["k"], complements=g["complements"], assume_simplitigs=True, want_maxone=True)
        assert r.ms.decode() == g["ms"], g
        assert md5(r.maxone + b"\n") == g["maxone_md5"]


@pytest.mark.parametrize("name", ["k31", "k31u", "k25", "k17u"])
def test_compute_S_simplitigs_md5(ctx, golden, simplitigs_bytes, name):
    g = golden["simplitigs_S"][name]
    k = int(name[1:].rstrip("u"))
    seq, off, ln = kb.frame_fasta(simplitigs_bytes)
    r = ctx.compute(seq, off, ln, k=k, complements=not name.endswith("u"), assume_simplitigs=True, want_maxone=True)
    assert r.length == g["length"] and md5(r.ms + b"\n") == g["md5"] and md5(r.maxone + b"\n") == g["maxone_md5"]


def test_compute_S_rejects_bad_records(ctx):
    seq, off, ln = orc.records_to_arrays([b"ACGTN", b"ACGTA"])
    with pytest.raises(kb.KcError) as e:
        ctx.compute(seq, off, ln, k=3, assume_simplitigs=True)
    assert e.value.code == -5
    seq, off, ln = orc.records_to_arrays([b"AC"])
    with pytest.raises(kb.KcError):
        ctx.compute(seq, off, ln, k=3, assume_simplitigs=True)


# ---- from-FASTA regime: set-exact, length / ones within 0.1 % of the reference ----------------------------------
@pytest.mark.parametrize("name", ["k31", "k13", "k63", "k127", "k31u", "k63u", "k127u", "k31z2", "k32", "k64u"])
def test_compute_spneumoniae(ctx, golden, spneumoniae_bytes, name):
    g = golden["spneumoniae_compute"][name]
    flags = name.lstrip("k0123456789")
    k = int(name[1:len(name) - len(flags)])
    compl = "u" not in flags
    z = 2 if "z2" in flags else 1
    seq, off, ln = kb.frame_fasta(spneumoniae_bytes)
    r = ctx.compute(seq, k=k, complements=compl, min_frequency=z, want_maxone=True)
    assert r.n_kmers == g["n_kmers"]
    want_k, want_v = orc.count_kmers(seq, off, ln, k, compl)
    want_k = want_k[want_v.astype(int) + 1 >= z]
    assert orc.verify_ms(r.ms, k, compl, want_k)        # verify.py: exactly the reference's k-mer set
    assert orc.verify_ms(r.maxone, k, compl, want_k)
    ones = sum(1 for c in r.ms if c <= 90)
    assert ones == g["ones"] == r.n_kmers                # min-one mask: every k-mer ON exactly once
    assert r.ms[-(k - 1):].islower() if k > 1 else True
    assert abs(r.length - g["length"]) <= 0.001 * g["length"]
    if "maxone_ones" in g:
        mo_ones = sum(1 for c in r.maxone if c <= 90)
        assert abs(mo_ones - g["maxone_ones"]) <= 0.001 * g["maxone_ones"]
    assert r.ms.upper() == r.maxone.upper()


def test_compute_errors(ctx):
    seq = np.frombuffer(b"ACGNNNAC\n", dtype=np.uint8)
    with pytest.raises(kb.KcError) as e:
        ctx.compute(seq, k=5)
    assert e.value.code == -4  # no k-mers (reference src/main.cpp:155-158)
    for bad in [dict(k=0), dict(k=128), dict(k=5, min_frequency=0), dict(k=5, min_frequency=256)]:
        with pytest.raises(kb.KcError) as e:
            ctx.compute(seq, **bad)
        assert e.value.code == -2


def test_compute_small_cases_vs_properties(ctx):
    """Tiny and degenerate inputs: single k-mer, palindromes (even k), homopolymers, circular sequence."""
    cases = [(b"ACG\n", 3), (b"ACGT\n", 4), (b"AAAAAAAA\n", 3), (b"ACGTAC\nGTACGT\n", 3), (b"ACGTACGTACGTACGT\n", 4),
             (b"ATATATATAT\n", 2), (b"ACGT\n", 1), (b"GATTACA\nTGTAATC\n", 5)]
    for text, k in cases:
        for compl in (True, False):
            seq = np.frombuffer(text, dtype=np.uint8)
            s2, off, ln = orc.records_to_arrays(text.split(b"\n")[:-1])
            want_k, _ = orc.count_kmers(s2, off, ln, k, compl)
            r = ctx.compute(seq, k=k, complements=compl, want_maxone=True)
            assert orc.verify_ms(r.ms, k, compl, want_k), (text, k, compl, r.ms)
            assert orc.verify_ms(r.maxone, k, compl, want_k)
            assert sum(1 for c in r.ms if c <= 90) == len(want_k)


def test_compute_device_matches_host_path(ctx, spneumoniae_bytes):
    import torch
    seq, _, _ = kb.frame_fasta(spneumoniae_bytes)
    host = ctx.compute(seq, k=31)
    d = torch.from_numpy(seq).cuda()
    torch.cuda.synchronize()
    dev = ctx.compute_device(d.data_ptr(), d.numel(), k=31)
    assert dev.length == host.length and dev.n_kmers == host.n_kmers
    assert ctx.copy_to_host(dev.ms_ptr, dev.length) == host.ms


# ---- BASELINE configs[1]: synthetic 50 Mbp, k = 31 — size-independent properties -----------------------------
def test_compute_config2_full_size(ctx):
    recs = synth.random_genome_records(50, 1_000_000, 12345)
    seq, off, ln = synth.frame_records(recs)
    r = ctx.compute(seq, k=31, want_maxone=True)
    assert r.n_kmers == 49_998_500                       # oracle U on this input
    assert abs(r.length - 49_999_860) <= 0.001 * 49_999_860
    got, n_on = orc.ms_kmers(r.ms, 31, True)
    assert n_on == r.n_kmers and len(got) == r.n_kmers   # every k-mer ON exactly once, all distinct
    keys, vals = ctx.count_kmers(seq, k=31)
    assert np.array_equal(got, keys)                      # superstring k-mers == counted set (checksum of sets)
    assert np.all(keys[1:, 0] > keys[:-1, 0])             # sortedness
    want_k, want_v = orc.count_kmers(seq, off, ln, 31, True)
    assert np.array_equal(keys, want_k) and np.array_equal(vals, want_v)
    assert r.ms.upper() == r.maxone.upper()


# ---- the multi-kernel level loop and the single-CTA small engine must agree ------------------------------------
def test_small_engine_and_host_levels_agree(ctx, golden, simplitigs_bytes):
    seq, off, ln = kb.frame_fasta(simplitigs_bytes)
    g = golden["simplitigs_S"]["k31"]
    try:
        ctx.set_option("small_engine", 0)
        for fz in golden["fuzz_S"]:
            s2, o2, l2 = orc.records_to_arrays([r.encode() for r in fz["records"]])
            r = ctx.compute(s2, o2, l2, k=fz["k"], complements=fz["complements"], assume_simplitigs=True, want_maxone=True)
            assert r.ms.decode() == fz["ms"] and md5(r.maxone + b"\n") == fz["maxone_md5"]
        r = ctx.compute(seq, off, ln, k=31, assume_simplitigs=True, want_maxone=True)
        assert md5(r.ms + b"\n") == g["md5"] and md5(r.maxone + b"\n") == g["maxone_md5"]
    finally:
        ctx.set_option("small_engine", 1)
    r = ctx.compute(seq, off, ln, k=31, assume_simplitigs=True, want_maxone=True)
    assert md5(r.ms + b"\n") == g["md5"] and md5(r.maxone + b"\n") == g["maxone_md5"]


# ---- histogram-free set construction (kmerset_fast.cuh) must agree with the exact one -----------------------------
def _fast_options(ctx, **kw):
    # any explicit option switches the signature-bucket construction (kmerset_sig.cuh, tried first by default) off, so that these
    # tests reach the fixed-slot / exact constructions; a bare call restores the defaults
    defaults = dict(fast_set=1, fast_leaf_target=768, fast_sigmas=8, fast_min_items=1 << 16, fast_heuristics=0, sig_set=0 if kw else 1)
    defaults.update(kw)
    for name, value in defaults.items():
        ctx.set_option(name, value)


@pytest.mark.parametrize("k,compl,z", [(31, True, 1), (21, False, 1), (31, True, 2), (47, True, 1), (95, False, 2), (127, True, 1)])
def test_fast_set_matches_exact(ctx, k, compl, z):
    recs = synth.random_genome_records(6, 50_000, 7 + k)
    recs.append(recs[0][1000:9000].copy())                      # duplicates
    recs.append(np.frombuffer(b"ACGTNNNNNACGT" * 50, dtype=np.uint8).copy())  # N breaks, low complexity
    recs += list(synth.reads_from_genome(20000, 4.0, 150, 0.01, seed=k))
    seq, off, ln = synth.frame_records(recs)
    default_ctas = ctx.stat("fast_max_ctas")
    try:
        _fast_options(ctx, fast_set=0)
        want = ctx.compute(seq, k=k, complements=compl, min_frequency=z)
        for leaf_target in (768, 16, 1):                         # 2, 2 and 3 partition levels on this input
            _fast_options(ctx, fast_min_items=0, fast_leaf_target=leaf_target)
            runs0, fb0 = ctx.stat("fast_runs"), ctx.stat("fast_fallbacks")
            got = ctx.compute(seq, k=k, complements=compl, min_frequency=z)
            assert ctx.stat("fast_runs") == runs0 + 1 and ctx.stat("fast_fallbacks") == fb0, leaf_target
            assert got.n_kmers == want.n_kmers and got.ms == want.ms, leaf_target
        # the grid size of the level >= 1 scatter (CTAs walk runs of consecutive tiles) never changes the result
        for leaf_target in (768, 16):
            _fast_options(ctx, fast_min_items=0, fast_leaf_target=leaf_target)
            for max_ctas in (1, 296, 4096, 1 << 20):
                ctx.set_option("fast_max_ctas", max_ctas)
                runs0, fb0 = ctx.stat("fast_runs"), ctx.stat("fast_fallbacks")
                got = ctx.compute(seq, k=k, complements=compl, min_frequency=z)
                assert ctx.stat("fast_runs") == runs0 + 1 and ctx.stat("fast_fallbacks") == fb0, (leaf_target, max_ctas)
                assert got.n_kmers == want.n_kmers and got.ms == want.ms, (leaf_target, max_ctas)
    finally:
        ctx.set_option("fast_max_ctas", default_ctas)
        _fast_options(ctx)
    want_k, want_v = orc.count_kmers(seq, off, ln, k, compl)
    want_k = want_k[want_v.astype(int) + 1 >= z]
    assert got.n_kmers == len(want_k) and orc.verify_ms(got.ms, k, compl, want_k)


def test_fast_set_single_level(ctx):
    recs = synth.random_genome_records(2, 50_000, 5)             # 100 K windows: one 8-bit level straight into the leaves
    seq, off, ln = synth.frame_records(recs)
    try:
        _fast_options(ctx, fast_set=0)
        want = ctx.compute(seq, k=31)
        _fast_options(ctx, fast_min_items=0)
        runs0 = ctx.stat("fast_runs")
        got = ctx.compute(seq, k=31)
        assert ctx.stat("fast_runs") == runs0 + 1
        assert got.ms == want.ms and got.n_kmers == want.n_kmers == 2 * (50_000 - 30)
    finally:
        _fast_options(ctx)


def test_fast_set_overflow_falls_back(ctx):
    recs = synth.random_genome_records(4, 60_000, 99)
    seq, off, ln = synth.frame_records(recs)
    try:
        _fast_options(ctx, fast_set=0)
        want = ctx.compute(seq, k=31)
        _fast_options(ctx, fast_min_items=0, fast_leaf_target=16, fast_sigmas=0)  # slots of exactly the mean: must overflow
        fb0 = ctx.stat("fast_fallbacks")
        got = ctx.compute(seq, k=31)
        assert ctx.stat("fast_fallbacks") == fb0 + 1
        assert got.ms == want.ms and got.n_kmers == want.n_kmers
        # one k-mer repeated far beyond any slot: a leaf overflows with the default plan
        text = b"A" * 300_000 + b"\n" + bytes(recs[0]) + b"\n"
        s2 = np.frombuffer(text, dtype=np.uint8)
        _fast_options(ctx, fast_set=0)
        want = ctx.compute(s2, k=31)
        _fast_options(ctx)
        fb0 = ctx.stat("fast_fallbacks")
        got = ctx.compute(s2, k=31)
        assert ctx.stat("fast_fallbacks") == fb0 + 1
        assert got.ms == want.ms and got.n_kmers == want.n_kmers
    finally:
        _fast_options(ctx)


def test_fast_set_heuristics(ctx):
    """-z > 1 announces a read set: the fixed-slot attempt is skipped; so is the call after an overflow on a similar input."""
    reads = synth.reads_from_genome(30_000, 10.0, 150, 0.01, seed=3)
    seq, off, ln = synth.frame_records(list(reads))
    try:
        _fast_options(ctx, fast_min_items=0, fast_heuristics=1)
        ctx.set_option("sig_set", 0)   # fast_heuristics resets the memory of both constructions
        runs0, fb0 = ctx.stat("fast_runs"), ctx.stat("fast_fallbacks")
        a = ctx.compute(seq, k=31, min_frequency=2)
        assert (ctx.stat("fast_runs"), ctx.stat("fast_fallbacks")) == (runs0, fb0)          # not even tried
        s2 = np.frombuffer(b"A" * 300_000 + b"\n" + bytes(reads[0]) + b"\n", dtype=np.uint8)   # one k-mer far beyond any slot
        b = ctx.compute(s2, k=31)                                                             # tried, overflows
        assert ctx.stat("fast_fallbacks") == fb0 + 1
        c = ctx.compute(s2, k=31)                                                             # remembered: not tried again
        assert (ctx.stat("fast_runs"), ctx.stat("fast_fallbacks")) == (runs0, fb0 + 1)
        assert b.ms == c.ms and b.n_kmers == c.n_kmers and a.n_kmers > 0
        d = ctx.compute(seq[:40_000], k=31)                                                   # a much smaller input: tried again
        assert ctx.stat("fast_runs") == runs0 + 1 and d.n_kmers > 0
    finally:
        _fast_options(ctx)
        ctx.set_option("fast_heuristics", 1)


def test_fast_set_config2_full_size(ctx):
    """configs[1] at full size through the path bench.py times (no -M): fast construction, exact set."""
    recs = synth.random_genome_records(50, 1_000_000, 12345)
    seq, off, ln = synth.frame_records(recs)
    runs0, fb0 = ctx.stat("fast_runs"), ctx.stat("fast_fallbacks")
    ctx.set_option("sig_set", 0)
    try:
        r = ctx.compute(seq, k=31)
    finally:
        ctx.set_option("sig_set", 1)
    assert ctx.stat("fast_runs") == runs0 + 1 and ctx.stat("fast_fallbacks") == fb0
    assert r.n_kmers == 49_998_500 and abs(r.length - 49_999_860) <= 0.001 * 49_999_860
    got, n_on = orc.ms_kmers(r.ms, 31, True)
    keys, _ = ctx.count_kmers(seq, k=31)
    assert n_on == r.n_kmers and np.array_equal(got, keys)


def test_compute_human_like_genome_with_repeats(ctx):
    """Scale model of configs[4] with the repeat model of SURVEY.md §8d (synth.human_like_genome: 24 records, ~5 % of the bases
    copied from earlier positions with 1-10 % divergence, half of the copies reverse-complemented, N runs): exact k-mer set,
    every k-mer ON exactly once, and more simplitigs than records (the repeats break the first-occurrence runs)."""
    recs = synth.human_like_genome(3_000_000, seed=3100)
    seq, off, ln = synth.frame_records(recs)
    want_k, _ = orc.count_kmers(seq, off, ln, 31, True)
    r = ctx.compute(seq, k=31)
    keys, _ = ctx.count_kmers(seq, k=31)
    assert np.array_equal(keys, want_k) and r.n_kmers == len(want_k)
    got, n_on = orc.ms_kmers(r.ms, 31, True)
    assert n_on == r.n_kmers and np.array_equal(got, keys)
    assert r.n_nodes > len(recs)
    assert r.length < r.n_kmers + 30 * r.n_nodes + 1        # every overlap the greedy found shortens the superstring

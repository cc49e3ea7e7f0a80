"""GPU: SURVEY.md §8 row a-8 — the sparse switch of src/main.cpp:94,175-181 and PartialPreSort (src/global_sparse.h:14-35)."""
import json
import os

import numpy as np
import pytest

import kmercamel_b200 as kb
from kmercamel_b200 import synth
from oracle import orc

from conftest import GOLDEN_DIR

pytestmark = pytest.mark.gpu
BIG = json.load(open(os.path.join(GOLDEN_DIR, "golden_big.json")))


def K(s, k=None):
    return orc.kmer_from_string(s, k)


def test_partial_presort_kats(ctx):
    """reference tests/global_sparse_unittest.h:11-41"""
    for kmers, k, want in [(["GTA", "TAC", "GGC"], 3, ["GGC", "GTA", "TAC"]),
                           (["TTTTTTTTTTTTT", "AAAAAAAAAAAAA", "GCGCGCGCGCGCG"], 13, ["AAAAAAAAAAAAA", "GCGCGCGCGCGCG", "TTTTTTTTTTTTT"]),
                           (["AAAAAAAAAAAAT", "AAAAAAAAAAAAA"], 13, ["AAAAAAAAAAAAT", "AAAAAAAAAAAAA"])]:
        got = ctx.partial_presort(np.stack([K(x, k) for x in kmers]), k=k)
        assert np.array_equal(got, np.stack([K(x, k) for x in want])), kmers


@pytest.mark.parametrize("k,n", [(1, 50), (2, 1000), (3, 5000), (13, 200_000), (31, 300_000), (40, 100_000), (70, 50_000), (127, 20_000)])
def test_partial_presort_random_vs_oracle(ctx, k, n):
    rng = np.random.default_rng(k * 1000 + 1)
    L = orc.limbs_for_k(k)
    kmers = rng.integers(0, 1 << 63, size=(n, L), dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=(n, L), dtype=np.uint64)
    for limb in range(L):                                # keep the low 2k bits of the word
        bits = min(64, max(0, 2 * k - 64 * limb))
        kmers[:, limb] &= np.uint64((1 << bits) - 1)
    kmers[: n // 3] = kmers[n // 3: 2 * (n // 3)]        # equal digits and equal words: stability matters
    got = ctx.partial_presort(kmers, k=k)
    assert np.array_equal(got, orc.partial_presort(kmers, k))


def test_global_sparse_strings(ctx):
    """reference tests/global_sparse_unittest.h:112-135 (GlobalSparse on explicit k-mer vectors): k-mer nodes in the given order =
    `-S` records of exactly k bases."""
    cases = [("TACgt", 3, ["CGT", "TAC", "ACG"], False),
             ("ACgTtt", 3, ["CGT", "TTT", "ACG"], False),
             ("TActt", 4, ["TACT", "ACTT"], False),
             ("TActTaaGgac", 4, ["TACT", "ACTT", "GGAC", "TAAG"], False),
             ("TTtcttttttttttttttttttttttttttga", 31, ["TTTCTTTTTTTTTTTTTTTTTTTTTTTTTTG", "TTCTTTTTTTTTTTTTTTTTTTTTTTTTTGA"], False),
             ("AtTTgtt", 4, ["ACAA", "ATTT", "AACA"], True)]
    for want, k, kmers, compl in cases:
        seq, off, ln = orc.records_to_arrays([x.encode() for x in kmers])
        r = ctx.compute(seq, off, ln, k=k, complements=compl, assume_simplitigs=True)
        assert r.ms == want.encode(), (kmers, r.ms)


@pytest.mark.parametrize("name", ["tiny_sparse_k31", "tiny_sparse_k9u"])
def test_sparse_switch_end_to_end(ctx, name):
    """An input whose simplitigs are nearly single k-mers: the greedy runs on the k-mers (n_nodes == n_kmers), the result
    represents exactly the reference's set and is as short as the reference's within 1 % (the reference's node order is its
    hash-table order, so ties break differently; with the switch off the superstring is measurably longer)."""
    g = BIG[name]
    ref = g["reference"]
    assert ref["sparse"] == 1
    seq, off, ln = synth.big_config_input(name)
    k, compl = g["k"], g["complements"]
    want_keys, _ = orc.count_kmers(seq, off, ln, k, compl)
    r = ctx.compute(seq, k=k, complements=compl)
    assert r.n_kmers == ref["n_kmers"] == len(want_keys)
    assert r.n_simplitigs * 5 >= r.n_kmers and r.n_nodes == r.n_kmers
    assert orc.verify_ms(r.ms, k, compl, want_keys)
    assert ctx.kmer_digest(r.ms, k=k, complements=compl, masked=True)[:3] == ref["digest"][:3]
    assert abs(r.length - ref["length"]) <= 0.01 * ref["length"], (r.length, ref["length"])
    ctx.set_option("sparse_switch", 0)
    try:
        r0 = ctx.compute(seq, k=k, complements=compl)
    finally:
        ctx.set_option("sparse_switch", 1)
    assert r0.n_nodes == r0.n_simplitigs and orc.verify_ms(r0.ms, k, compl, want_keys)
    lb, _ = ctx.lower_bound(seq, k=k, complements=compl)
    assert lb <= r.length

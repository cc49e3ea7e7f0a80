"""CPU: the oracle's k-mer digest against the reference's (tests/golden/golden_big.json, written by make_golden_big.py from
oracle/_ref/ref_harness `full`) on the tiny pinned configurations, all three word widths; and the seeded generators."""
import json
import os
import zlib

import numpy as np
import pytest

from kmercamel_b200 import synth
from oracle import orc

from conftest import GOLDEN_DIR

BIG = json.load(open(os.path.join(GOLDEN_DIR, "golden_big.json")))


@pytest.mark.parametrize("name", ["tiny_k31z2", "tiny_k63", "tiny_k127u"])
def test_oracle_digest_matches_reference(name):
    g = BIG[name]
    seq, off, ln = synth.big_config_input(name)
    assert len(seq) == g["n_bytes"] and zlib.crc32(seq.tobytes()) == g["crc32"]
    keys, vals = orc.count_kmers(seq, off, ln, g["k"], g["complements"])
    d = orc.kmer_digest(keys, vals, g["min_frequency"])
    assert d == g["reference"]["digest"]
    assert d[0] == g["reference"]["n_kmers"] == g["reference"]["ones"]


def test_reads_chunks_model():
    """reads_chunks: chunks are reproducible on their own, reads are genome substrings up to ~1 % substitutions."""
    a = list(synth.reads_chunks(50_000, 4.0, 150, 0.01, 5, chunk_reads=400))
    b = list(synth.reads_chunks(50_000, 4.0, 150, 0.01, 5, chunk_reads=400))
    assert len(a) == 4 and all(np.array_equal(x, y) for x, y in zip(a, b))
    clean = list(synth.reads_chunks(50_000, 4.0, 150, 0.0, 5, chunk_reads=400))
    diff = sum(int((x != y).sum()) for x, y in zip(a, clean))
    total = sum(x.size for x in a)
    assert 0.005 < diff / total < 0.015
    genome = synth._ACGT[np.random.default_rng([5, 0]).integers(0, 4, size=50_000, dtype=np.uint8)].tobytes()
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    for r in clean[0][:20]:
        s = r.tobytes()
        assert s in genome or s.translate(comp)[::-1] in genome


def test_golden_big_records_are_consistent():
    for name, g in BIG.items():
        r = g["reference"]
        assert r["digest"][0] == r["n_kmers"]
        if "length" in r:
            assert r["ones"] == r["n_kmers"] and r["tail_lower"] and r["length"] >= r["n_kmers"] + g["k"] - 1


def test_partial_presort_kats():
    """reference tests/global_sparse_unittest.h:11-41"""
    K = orc.kmer_from_string
    for kmers, k, want in [(["GTA", "TAC", "GGC"], 3, ["GGC", "GTA", "TAC"]),
                           (["TTTTTTTTTTTTT", "AAAAAAAAAAAAA", "GCGCGCGCGCGCG"], 13, ["AAAAAAAAAAAAA", "GCGCGCGCGCGCG", "TTTTTTTTTTTTT"]),
                           (["AAAAAAAAAAAAT", "AAAAAAAAAAAAA"], 13, ["AAAAAAAAAAAAT", "AAAAAAAAAAAAA"])]:
        got = orc.partial_presort(np.stack([K(x, k) for x in kmers]), k)
        assert np.array_equal(got, np.stack([K(x, k) for x in want]))

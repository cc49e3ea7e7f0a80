#!/usr/bin/env python3
"""Regenerates tests/golden/golden.json from the UNMODIFIED reference compiled into oracle/_ref
(`make -C oracle ref`).  Runs only in the build container (needs /root/reference); the JSON and the fixture
files next to it are what travels to the GPU box.

Fixtures copied verbatim from the reference's data/test directories (inputs, not source code):
  spneumoniae.fa.gz   <- data/spneumoniae.fa          (BASELINE.json configs[0])
  simplitigs-k31.fa.gz<- data/simplitigs-k31.fa
  test.fa             <- tests/testdata/test.fa       (parser golden vectors, tests/parser_unittest.h:38-91)
"""
import gzip
import hashlib
import json
import os
import random
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref", "kmercamel")
HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_harness")


def md5(b: bytes) -> str:
    return hashlib.md5(b).hexdigest()


def run_ref(path, k, flags=(), want_maxone=False):
    """-> dict(length, ones, md5 of line 2 incl. newline [, maxone_*])"""
    with tempfile.TemporaryDirectory() as td:
        args = [REF, "compute", "-k", str(k), *flags]
        mo = os.path.join(td, "mo.fa")
        if want_maxone:
            args += ["-M", mo]
        args.append(path)
        p = subprocess.run(args, capture_output=True)
        assert p.returncode == 0, p.stderr
        line2 = p.stdout.split(b"\n")[1]
        out = dict(length=len(line2), ones=sum(1 for c in line2 if c <= 90), md5=md5(line2 + b"\n"))
        log = p.stderr.decode()
        for ln in log.splitlines():
            if "Finished collecting k-mers:" in ln:
                out["n_kmers"] = int(ln.split("k-mers:")[1].split()[0])
            if "simplitigs (" in ln:
                out["n_simplitigs"] = int(ln.split("simplitigs (")[1].split()[0])
        if want_maxone:
            l2 = open(mo, "rb").read().split(b"\n")[1]
            out["maxone_ones"] = sum(1 for c in l2 if c <= 90)
            out["maxone_md5"] = md5(l2 + b"\n")
        return out, line2


def run_ref_lowerbound(path, k, flags=()):
    """`kmercamel lowerbound` of the reference (src/main.cpp:378-443): the integer it prints."""
    p = subprocess.run([REF, "lowerbound", "-k", str(k), *flags, path], capture_output=True)
    assert p.returncode == 0, p.stderr
    return int(p.stdout.split()[0])


def add_lowerbound(G, td, sp, st):
    """Section "lowerbound": -S inputs are exact (pure function of record order); from-FASTA values depend on the
    reference's khash-order simplitigs and are compared with a tolerance."""
    L = dict(fuzz_S=[], simplitigs_S={}, spneumoniae={})
    for g in G["fuzz_S"]:
        f = os.path.join(td, "fzlb.fa")
        open(f, "w").write("".join(f">{i}\n{r}\n" for i, r in enumerate(g["records"])))
        L["fuzz_S"].append(run_ref_lowerbound(f, g["k"], ("-S",) + (() if g["complements"] else ("-u",))))
    for name, k, flags in [("k31", 31, ("-S",)), ("k31u", 31, ("-S", "-u")), ("k25", 25, ("-S",)), ("k17u", 17, ("-S", "-u"))]:
        L["simplitigs_S"][name] = run_ref_lowerbound(st, k, flags)
    for name, k, flags in [("k31", 31, ()), ("k31u", 31, ("-u",)), ("k13", 13, ()), ("k63", 63, ()), ("k31z2", 31, ("-z", "2"))]:
        L["spneumoniae"][name] = run_ref_lowerbound(sp, k, flags)
    G["lowerbound"] = L


def kmers_dump(path, k, complements):
    with tempfile.TemporaryDirectory() as td:
        o = os.path.join(td, "k.bin")
        subprocess.check_call([HARNESS, "kmers", path, str(k), str(int(complements)), o])
        raw = open(o, "rb").read()
    n = int.from_bytes(raw[:8], "little")
    limbs = int.from_bytes(raw[8:12], "little")
    keys = raw[12:12 + n * limbs * 8]
    vals = raw[12 + n * limbs * 8:]
    hist = {}
    for v in vals:
        hist[v] = hist.get(v, 0) + 1
    return dict(n=n, limbs=limbs, keys_md5=md5(keys), vals_md5=md5(vals),
                hist={str(a): b for a, b in sorted(hist.items())[:8]})


def rc(s):
    return s[::-1].translate(str.maketrans("ACGT", "TGCA"))


def fuzz_records(rng, k, n, alpha):
    g = "".join(rng.choice(alpha) for _ in range(rng.randint(k + 5, 4 * n + k + 5)))
    recs = []
    for _ in range(n):
        t = rng.random()
        if t < 0.5 and len(g) > k:
            a = rng.randrange(0, len(g) - k)
            ln = rng.randint(k, min(len(g) - a, k + rng.randint(0, 12)))
            s = g[a:a + ln]
        elif t < 0.7 and recs:
            s = rng.choice(recs)
        elif t < 0.85 and recs:
            s = rc(rng.choice(recs))
        else:
            s = "".join(rng.choice(alpha) for _ in range(rng.randint(k, k + 6)))
        recs.append(s)
    return recs


def main():
    import sys
    G = {}
    with tempfile.TemporaryDirectory() as td:
        sp = os.path.join(td, "spneumoniae.fa")
        open(sp, "wb").write(gzip.open(os.path.join(HERE, "spneumoniae.fa.gz")).read())
        st = os.path.join(td, "simplitigs-k31.fa")
        open(st, "wb").write(gzip.open(os.path.join(HERE, "simplitigs-k31.fa.gz")).read())
        if "--only-lowerbound" in sys.argv:  # extend the committed JSON without re-running the other sections
            G = json.load(open(os.path.join(HERE, "golden.json")))
            add_lowerbound(G, td, sp, st)
            json.dump(G, open(os.path.join(HERE, "golden.json"), "w"), indent=1, sort_keys=True)
            print("lowerbound:", G["lowerbound"]["simplitigs_S"], G["lowerbound"]["spneumoniae"])
            return
        # from-FASTA regime: sizes/lengths (output bytes depend on khash order -> compared by set + tolerance)
        G["spneumoniae_compute"] = {}
        for name, k, flags, mo in [("k31", 31, (), True), ("k13", 13, (), False), ("k63", 63, (), False),
                                   ("k127", 127, (), False), ("k31u", 31, ("-u",), True), ("k63u", 63, ("-u",), False),
                                   ("k127u", 127, ("-u",), False), ("k31z2", 31, ("-z", "2"), False),
                                   ("k32", 32, (), False), ("k64u", 64, ("-u",), False)]:
            G["spneumoniae_compute"][name], _ = run_ref(sp, k, flags, mo)
        # -S regime: byte-exact
        G["simplitigs_S"] = {}
        for name, k, flags in [("k31", 31, ("-S",)), ("k31u", 31, ("-S", "-u")), ("k25", 25, ("-S",)), ("k17u", 17, ("-S", "-u"))]:
            G["simplitigs_S"][name], _ = run_ref(st, k, flags, True)
        # stage 1: sorted k-mer set + uint8 values
        G["spneumoniae_kmers"] = {}
        for name, k, c in [("k31", 31, True), ("k31u", 31, False), ("k63", 63, True), ("k127u", 127, False),
                           ("k32", 32, True), ("k64", 64, True), ("k5", 5, True), ("k1u", 1, False)]:
            G["spneumoniae_kmers"][name] = kmers_dump(sp, k, c)
        G["test_fa_kmers"] = {}
        tf = os.path.join(HERE, "test.fa")
        for name, k, c in [("k3", 3, True), ("k3u", 3, False), ("k10", 10, True), ("k5", 5, True), ("k5u", 5, False),
                           ("k2", 2, True), ("k2u", 2, False), ("k1u", 1, False), ("k4", 4, True)]:
            G["test_fa_kmers"][name] = kmers_dump(tf, k, c)
        # parser edge cases: text -> number of distinct k-mers and the superstring
        cases = {
            "plain": ">a\nACGTAC\n", "crlf": ">a\r\nACGTAC\r\n", "lower": ">a\nacgtac\n", "nonl": ">a\nACGTAC",
            "blank": ">a\nACG\n\nTAC\n", "junk": "junk\n>a\nACGTAC\n", "gt_mid": ">a\nACG>TAC\n", "space": ">a\nACG TAC\n",
            "n": ">a\nACGNTAC\n", "two": ">a\nACG\n>b\nTAC\n", "fastq": "@r1\nACGTAC\n+\nIIIIII\n@r2\nTTTGA\n+\n@>III\n",
            "fastq_trunc": "@r1\nACGTAC\n+\nIII\n", "comment": ">a some comment\nACGTAC\n", "tabname": ">a\tx\nACGTAC\n",
            "multiline": ">a\nAC\nGT\nAC\n", "empty_rec": ">a\n>b\nACGTAC\n", "iupac": ">a\nACGRYTAC\n",
        }
        G["parser_cases"] = {}
        for name, text in cases.items():
            f = os.path.join(td, name + ".fa")
            open(f, "w", newline="").write(text)
            d = kmers_dump(f, 3, False)
            try:
                res, line2 = run_ref(f, 3, ("-u",))
                ms = line2.decode()
            except AssertionError:
                ms = None
            G["parser_cases"][name] = dict(text=text, n_kmers_k3u=d["n"], keys_md5=d["keys_md5"], ms_k3u=ms)
        # fuzzed -S instances, byte-exact superstrings from the reference CLI
        rng = random.Random(20261017)
        G["fuzz_S"] = []
        for it in range(60):
            k = rng.choice([2, 3, 4, 5, 7, 11, 15, 31, 32, 40, 70, 127])
            n = rng.randint(1, 40 if k < 100 else 12)
            alpha = rng.choice(["ACGT", "AC", "ACG", "ACGT"])
            compl = rng.random() < 0.6
            recs = fuzz_records(rng, k, n, alpha)
            f = os.path.join(td, "fz.fa")
            open(f, "w").write("".join(f">{i}\n{r}\n" for i, r in enumerate(recs)))
            res, line2 = run_ref(f, k, ("-S",) + (() if compl else ("-u",)), True)
            mo = None
            G["fuzz_S"].append(dict(k=k, complements=compl, records=recs, ms=line2.decode(), maxone_md5=res["maxone_md5"]))
        add_lowerbound(G, td, sp, st)
    json.dump(G, open(os.path.join(HERE, "golden.json"), "w"), indent=1, sort_keys=True)
    print("wrote golden.json:", {k: len(v) for k, v in G.items()})


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""Regenerates tests/golden/golden_masks.json from the UNMODIFIED reference CLI compiled into oracle/_ref
(`make -C oracle ref`): `compute -a streaming [-z]`, `maskopt -t max-one|min-one` and the four text conversions
(SURVEY.md §8f rows 2-4).  Runs only in the build container (needs /root/reference); the JSON travels to the GPU box.

Inputs are either the committed fixtures (spneumoniae.fa.gz, simplitigs-k31.fa.gz, test.fa), seeded synthetic reads
(kmercamel_b200.synth, regenerated identically by the tests) or small random cases stored in full.
"""
import gzip
import hashlib
import json
import os
import random
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.path.join(ROOT, "oracle", "_ref", "kmercamel")

from kmercamel_b200 import synth  # noqa: E402  (numpy only)


def md5(b: bytes) -> str:
    return hashlib.md5(b).hexdigest()


def ref(args, stdin=None):
    p = subprocess.run([REF, *args], capture_output=True, input=stdin)
    assert p.returncode == 0, (args, p.stderr[-500:])
    return p.stdout


def line2(out: bytes) -> bytes:
    return out.split(b"\n")[1]


def summary(seq: bytes) -> dict:
    return dict(length=len(seq), ones=sum(1 for c in seq if c <= 90), md5=md5(seq))


def reads_fasta(genome_len, coverage, read_len, err, seed) -> bytes:
    reads = synth.reads_from_genome(genome_len, coverage, read_len, err, seed)
    return b"".join(b">r%d\n%s\n" % (i, r.tobytes()) for i, r in enumerate(reads))


READS = dict(genome_len=20000, coverage=6.0, read_len=100, err=0.01, seed=77)


def main():
    random.seed(20261017)
    g = {}
    with tempfile.TemporaryDirectory() as td:
        def tmp(name, data):
            path = os.path.join(td, name)
            open(path, "wb").write(data)
            return path

        sp = tmp("sp.fa", gzip.open(os.path.join(HERE, "spneumoniae.fa.gz")).read())
        sim = tmp("sim.fa", gzip.open(os.path.join(HERE, "simplitigs-k31.fa.gz")).read())
        tfa = os.path.join(HERE, "test.fa")
        reads = tmp("reads.fa", reads_fasta(**READS))

        # ---- streaming ----
        st = {}
        for name, path, k, flags in [("sp_k31", sp, 31, []), ("sp_k31u", sp, 31, ["-u"]), ("sp_k13", sp, 13, []), ("sp_k63", sp, 63, []),
                                     ("sp_k127u", sp, 127, ["-u"]), ("sp_k31z2", sp, 31, ["-z", "2"]), ("sp_k5z3", sp, 5, ["-z", "3"]),
                                     ("reads_k31", reads, 31, []), ("reads_k31z2", reads, 31, ["-z", "2"]), ("reads_k21z3u", reads, 21, ["-z", "3", "-u"]),
                                     ("reads_k47z2", reads, 47, ["-z", "2"]), ("sim_k31", sim, 31, [])]:
            out = ref(["compute", "-a", "streaming", "-k", str(k), *flags, path])
            st[name] = summary(line2(out))
            st[name]["header_tail"] = out.split(b"\n")[0].decode().split("' ", 1)[1]
        for k in (1, 2, 3, 4, 5):
            for flags in ([], ["-u"], ["-z", "2"], ["-z", "3", "-u"]):
                out = ref(["compute", "-a", "streaming", "-k", str(k), *flags, tfa])
                st["test_k%d%s" % (k, "".join(flags).replace("-", ""))] = dict(seq=line2(out).decode())
        g["streaming"] = st
        g["reads_params"] = READS

        # ---- streaming / maskopt fuzz: small random inputs stored in full ----
        fz = []
        for i in range(40):
            k = random.choice([1, 2, 3, 4, 5, 7, 9])
            n_rec = random.randint(1, 4)
            alphabet = "ACGT" * 6 + "acgt" * 2 + "N"
            recs = ["".join(random.choice(alphabet) for _ in range(random.randint(0, 60))) for _ in range(n_rec)]
            fasta = "".join(">r%d\n%s\n" % (j, r) for j, r in enumerate(recs)).encode()
            path = tmp("fz.fa", fasta)
            u = random.random() < 0.4
            z = random.choice([1, 1, 2, 3])
            flags = (["-u"] if u else []) + (["-z", str(z)] if z > 1 else [])
            out = ref(["compute", "-a", "streaming", "-k", str(k), *flags, path])
            fz.append(dict(k=k, unidirectional=u, z=z, records=recs, seq=line2(out).decode()))
        g["streaming_fuzz"] = fz

        mo = []
        for i in range(60):
            k = random.choice([1, 2, 3, 4, 5, 8, 11])
            n = random.randint(max(k, 1), 90)
            # low-complexity letters so that k-mers repeat; random case = random input mask
            base = "".join(random.choice("ACGT" if random.random() < 0.7 else "AC") for _ in range(n))
            ms = "".join(c if random.random() < 0.45 else c.lower() for c in base)
            path = tmp("mo.fa", (">superstring some comment\n%s\n" % ms).encode())
            u = random.random() < 0.4
            row = dict(k=k, unidirectional=u, ms=ms)
            for t in ("max-one", "min-one"):
                out = ref(["maskopt", "-t", t, "-k", str(k), *(["-u"] if u else []), path])
                row[t] = line2(out).decode()
                row["header_" + t] = out.split(b"\n")[0].decode()
            mo.append(row)
        g["maskopt_fuzz"] = mo

        # ---- maskopt on large inputs ----
        big = {}
        sim_ms_out = ref(["spss2ms", "-k", "31", sim])
        sim_ms = tmp("sim_ms.fa", sim_ms_out)
        big["sim_spss2ms_k31"] = summary(line2(sim_ms_out))
        for name, k, flags in [("k31", 31, []), ("k31u", 31, ["-u"]), ("k25", 25, []), ("k40u", 40, ["-u"])]:
            for t in ("max-one", "min-one"):
                out = ref(["maskopt", "-t", t, "-k", str(k), *flags, sim_ms])
                big["sim_%s_%s" % (name, t)] = summary(line2(out))
        # the streaming superstring of spneumoniae (many k-mers occur again in lower case) re-optimised
        sp_stream = tmp("sp_stream.fa", ref(["compute", "-a", "streaming", "-k", "31", sp]))
        for t in ("max-one", "min-one"):
            big["sp_stream_k31_" + t] = summary(line2(ref(["maskopt", "-t", t, "-k", "31", sp_stream])))
        sp_stream_u = tmp("sp_stream_u.fa", ref(["compute", "-a", "streaming", "-k", "70", "-u", sp]))
        for t in ("max-one", "min-one"):
            big["sp_stream_k70u_" + t] = summary(line2(ref(["maskopt", "-t", t, "-k", "70", "-u", sp_stream_u])))
        g["maskopt"] = big

        # ---- conversions: full texts of small random cases ----
        cv = []
        for i in range(30):
            k = random.choice([1, 2, 3, 5, 8])
            n = random.randint(0, 70)
            ms = "".join(random.choice("ACGTacgt") for _ in range(n))
            path = tmp("cv.fa", (">ms\n%s\n" % ms).encode())
            mpath, spath = os.path.join(td, "m.txt"), os.path.join(td, "s.txt")
            ref(["ms2mssep", "-m", mpath, "-s", spath, path])
            mask, sup = open(mpath).read(), open(spath).read()
            joined = ref(["mssep2ms", "-m", mpath, "-s", spath]).decode()
            spss = ref(["ms2spss", "-k", str(k), path]).decode()
            recs = ["".join(random.choice("ACGTacgt") for _ in range(random.randint(0, 25))) for _ in range(random.randint(1, 5))]
            rpath = tmp("rec.fa", "".join(">x%d\n%s\n" % (j, r) for j, r in enumerate(recs)).encode())
            back = line2(ref(["spss2ms", "-k", str(k), rpath])).decode()
            cv.append(dict(k=k, ms=ms, mask=mask, superstring=sup, joined=joined, spss=spss, records=recs, spss2ms=back))
        g["conversions"] = cv
    json.dump(g, open(os.path.join(HERE, "golden_masks.json"), "w"), indent=1, sort_keys=True)
    print("wrote golden_masks.json:", {k: len(v) for k, v in g.items()})


if __name__ == "__main__":
    main()

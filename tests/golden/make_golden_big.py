#!/usr/bin/env python3
"""Reference results for the FULL-SIZE BASELINE.json configurations -> tests/golden/golden_big.json.

Runs only in the build container (needs /root/reference compiled into oracle/_ref by `make -C oracle ref`): every input is
generated from a seed by kmercamel_b200/synth.py, written as a FASTA file under /tmp/kc_big, and pushed through
`oracle/_ref/ref_harness full` — the UNMODIFIED reference's ReadKMers[Filtered] -> get_simplitigs -> Global / GlobalSparse
chain (src/main.cpp:146-186) in one process.  Recorded per configuration: the generator call, crc32 of the framed sequence
(so the GPU box can tell that it regenerated the same input), n_kmers, the order-independent digest of the kept
(k-mer, count) pairs, the number of simplitigs, the superstring length and its number of ones, and the reference's stage
times on this container's host core.  The GPU tests (tests/test_gpu_big.py) regenerate the inputs from the same seeds and
compare kc_kmer_digest / kc_compute against these numbers; nothing under /root/reference is read at test time.

usage: make_golden_big.py NAME [NAME ...]      (names: see CONFIGS)
"""
import fcntl
import json
import os
import subprocess
import sys
import time
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from kmercamel_b200 import synth  # noqa: E402

HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
OUT = os.path.join(HERE, "golden_big.json")
TMP = "/tmp/kc_big"

# name -> (generator description, k, complements, min_frequency).  `make_input(name)` below is the single source of truth
# shared with tests/test_gpu_big.py through synth.big_config_input.
CONFIGS = synth.BIG_CONFIGS


def crc32_of(seq: np.ndarray) -> int:
    c = 0
    mv = memoryview(seq)
    for lo in range(0, len(seq), 1 << 28):
        c = zlib.crc32(mv[lo:lo + (1 << 28)], c)
    return c


def write_fasta(path: str, seq: np.ndarray, off: np.ndarray, ln: np.ndarray, one_line: bool):
    with open(path, "wb") as f:
        if one_line:                                  # reads: ">i\nREAD\n", written in blocks
            step = 1 << 18
            for lo in range(0, len(off), step):
                hi = min(len(off), lo + step)
                parts = []
                for i in range(lo, hi):
                    parts.append(b">%d\n" % i)
                    parts.append(seq[int(off[i]):int(off[i]) + int(ln[i]) + 1].tobytes())
                f.write(b"".join(parts))
        else:
            for i in range(len(off)):
                f.write(synth.fasta_bytes([seq[int(off[i]):int(off[i]) + int(ln[i])]], prefix=f"r{i}_"))


def main(names):
    os.makedirs(TMP, exist_ok=True)
    for name in names:
        cfg = CONFIGS[name]
        t0 = time.time()
        seq, off, ln = synth.big_config_input(name)
        rec = dict(config=cfg["config"], generator=cfg["generator"], k=cfg["k"], complements=cfg["complements"],
                   min_frequency=cfg["min_frequency"], n_bytes=int(len(seq)), n_records=int(len(off)), crc32=crc32_of(seq))
        fa = os.path.join(TMP, name + ".fa")
        write_fasta(fa, seq, off, ln, cfg.get("one_line", False))
        del seq
        rec["generate_s"] = round(time.time() - t0, 1)
        ms = os.path.join(TMP, name + ".ms")
        p = subprocess.run([HARNESS, "full", fa, str(cfg["k"]), str(int(cfg["complements"])), str(cfg["min_frequency"]), ms],
                           capture_output=True)
        assert p.returncode == 0, p.stderr[-2000:]
        ref = json.loads(p.stdout.decode().strip().splitlines()[-1])
        line = np.fromfile(ms, dtype=np.uint8)
        assert line[-1] == 10
        line = line[:-1]
        ref["length"] = int(len(line))
        ref["ones"] = int((line <= 90).sum())
        ref["tail_lower"] = bool((line[len(line) - (cfg["k"] - 1):] > 90).all())
        rec["reference"] = ref
        rec["reference_log"] = p.stderr.decode().splitlines()[-3:]
        os.remove(fa)
        os.remove(ms)
        with open(OUT + ".lock", "w") as lk:          # two lanes of this script may run side by side
            fcntl.flock(lk, fcntl.LOCK_EX)
            allr = json.load(open(OUT)) if os.path.exists(OUT) else {}
            allr[name] = rec
            json.dump(allr, open(OUT, "w"), indent=1, sort_keys=True)
        print(name, json.dumps(rec), flush=True)


if __name__ == "__main__":
    main(sys.argv[1:])

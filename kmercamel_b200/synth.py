"""Seeded synthetic inputs for the BASELINE.json configurations.  numpy only."""
from __future__ import annotations

import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def random_genome_records(n_records: int, record_len: int, seed: int):
    """n_records i.i.d. uniform ACGT records (config 2: 50 x 1,000,000, seed 12345)."""
    rng = np.random.default_rng(seed)
    return [_ACGT[rng.integers(0, 4, size=record_len, dtype=np.uint8)] for _ in range(n_records)]


def frame_records(records):
    """Records -> (seq, rec_off, rec_len) in the layout of kc_input (each record followed by '\\n')."""
    total = sum(len(r) + 1 for r in records)
    seq = np.empty(total, dtype=np.uint8)
    off = np.zeros(len(records), dtype=np.uint64)
    ln = np.zeros(len(records), dtype=np.uint64)
    p = 0
    for i, r in enumerate(records):
        off[i] = p
        ln[i] = len(r)
        seq[p:p + len(r)] = r
        seq[p + len(r)] = 10
        p += len(r) + 1
    return seq, off, ln


def fasta_bytes(records, width: int = 80, prefix: str = "r") -> bytes:
    """Multi-FASTA text with `width`-column lines and headers >r<i>."""
    parts = []
    for i, r in enumerate(records):
        parts.append(f">{prefix}{i}\n".encode())
        r = np.asarray(r, dtype=np.uint8)
        full = len(r) // width * width
        if full:
            body = np.empty((full // width, width + 1), dtype=np.uint8)
            body[:, :width] = r[:full].reshape(-1, width)
            body[:, width] = 10
            parts.append(body.tobytes())
        if full < len(r):
            parts.append(r[full:].tobytes() + b"\n")
    return b"".join(parts)


def reads_from_genome(genome_len: int, coverage: float, read_len: int, error_rate: float, seed: int):
    """Config 4: reads with uniform starts, 50 % reverse-complemented, i.i.d. substitution errors."""
    rng = np.random.default_rng(seed)
    codes = rng.integers(0, 4, size=genome_len, dtype=np.uint8)
    n_reads = int(genome_len * coverage / read_len)
    starts = rng.integers(0, genome_len - read_len + 1, size=n_reads)
    idx = starts[:, None] + np.arange(read_len)[None, :]
    reads = codes[idx]
    flip = rng.random(n_reads) < 0.5
    reads[flip] = (3 - reads[flip])[:, ::-1]
    err = rng.random(reads.shape) < error_rate
    reads[err] = (reads[err] + rng.integers(1, 4, size=int(err.sum()), dtype=np.uint8)) % 4
    return _ACGT[reads]  # [n_reads, read_len] uint8


# GRCh38 chromosome lengths in Mbp (1..22, X, Y): only their proportions are used.
_HUMAN_MBP = (248, 242, 198, 190, 182, 171, 159, 145, 138, 134, 135, 133, 114, 107, 102, 90, 83, 80, 59, 64, 47, 51, 156, 57)
_COMP = np.zeros(256, dtype=np.uint8)
for _a, _b in zip(b"ACGTN", b"TGCAN"):
    _COMP[_a] = _b


def human_like_genome(total_len: int, seed: int = 3100, repeat_fraction: float = 0.05, n_runs: int = 4):
    """Config 5 (BASELINE.json configs[4]): 24 records with human-like relative chromosome lengths summing to
    total_len, i.i.d. uniform ACGT, plus a repeat model so that the number of simplitigs is non-trivial:
      * about repeat_fraction of the bases are copies of EARLIER positions of the genome, in segments of 300 bp - 6 kbp, half
        of them reverse-complemented, each with its own divergence of 1 - 10 % i.i.d. substitutions (copies of copies occur);
      * n_runs runs of 'N' (1 - 50 kbp, at most total_len / 200 each).
    The pure-uniform variant of the same size is random_genome_records.  -> list of uint8 records."""
    rng = np.random.default_rng(seed)
    g = _ACGT[rng.integers(0, 4, size=total_len, dtype=np.uint8)]
    lo, hi = 300, 6000
    n_seg = int(repeat_fraction * total_len / ((lo + hi) / 2))
    for _ in range(n_seg):
        ln = int(rng.integers(lo, hi + 1))
        if total_len < 4 * ln:
            continue
        dst = int(rng.integers(ln, total_len - ln))
        src = int(rng.integers(0, dst - ln + 1))                       # an earlier segment
        seg = g[src:src + ln].copy()
        if rng.random() < 0.5:
            seg = _COMP[seg][::-1].copy()
        div = rng.uniform(0.01, 0.10)
        hit = np.flatnonzero(rng.random(ln) < div)
        if hit.size:                                                    # substitute by one of the three OTHER letters
            code = np.searchsorted(_ACGT, seg[hit])
            seg[hit] = _ACGT[(code + rng.integers(1, 4, size=hit.size)) % 4]
        g[dst:dst + ln] = seg
    for _ in range(n_runs):
        ln = int(min(rng.integers(1000, 50001), max(total_len // 200, 1)))
        at = int(rng.integers(0, total_len - ln + 1))
        g[at:at + ln] = ord("N")
    w = np.asarray(_HUMAN_MBP, dtype=np.float64)
    cuts = np.floor(np.cumsum(w) / w.sum() * total_len).astype(np.int64)
    cuts[-1] = total_len
    out, p = [], 0
    for c in cuts:
        if c > p:
            out.append(g[p:c].copy())
        p = int(c)
    return out

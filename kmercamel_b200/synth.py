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

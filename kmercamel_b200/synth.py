"""Seeded synthetic inputs for the BASELINE.json configurations.  numpy only."""
from __future__ import annotations

import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def random_genome_records(n_records: int, record_len: int, seed: int):
    """n_records i.i.d. uniform ACGT records (config 2: 50 x 1,000,000, seed 12345)."""
    rng = np.random.default_rng(seed)
    return [_ACGT[rng.integers(0, 4, size=record_len, dtype=np.uint8)] for _ in range(n_records)]


def short_records(n_records: int, min_len: int, max_len: int, seed: int):
    """Records barely longer than k: their simplitigs are nearly single k-mers, which makes the reference hand the individual
    k-mers to the greedy (5 * simplitigs >= k-mers, src/main.cpp:94,175-181)."""
    rng = np.random.default_rng(seed)
    lens = rng.integers(min_len, max_len + 1, size=n_records)
    return [_ACGT[rng.integers(0, 4, size=int(l), dtype=np.uint8)] for l in lens]


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


def reads_chunks(genome_len: int, coverage: float, read_len: int, error_rate: float, seed: int, chunk_reads: int = 1 << 20):
    """Config 4 at full size without the [n_reads, read_len] index matrix of reads_from_genome: yields uint8 arrays
    [n, read_len] of ACGT letters, chunk c drawn from default_rng([seed, c + 1]) (the genome from default_rng([seed, 0])), so
    any chunk can be regenerated on its own.  Same model: uniform starts, 50 % reverse-complemented, i.i.d. substitution
    errors — drawn as geometric gaps between error positions, which IS the Bernoulli(error_rate) process."""
    codes = np.random.default_rng([seed, 0]).integers(0, 4, size=genome_len, dtype=np.uint8)
    n_reads = int(genome_len * coverage / read_len)
    ar = np.arange(read_len, dtype=np.int64)
    for c, lo in enumerate(range(0, n_reads, chunk_reads)):
        n = min(chunk_reads, n_reads - lo)
        rng = np.random.default_rng([seed, c + 1])
        starts = rng.integers(0, genome_len - read_len + 1, size=n)
        reads = codes[starts[:, None] + ar[None, :]]
        flip = rng.random(n) < 0.5
        reads[flip] = (3 - reads[flip])[:, ::-1]
        if error_rate > 0:
            total, flat, at = n * read_len, reads.reshape(-1), -1
            while at < total:
                gaps = rng.geometric(error_rate, size=int(total * error_rate * 1.1) + 64)
                pos = at + np.cumsum(gaps)
                at = int(pos[-1])
                pos = pos[pos < total]
                flat[pos] = (flat[pos] + rng.integers(1, 4, size=pos.size, dtype=np.uint8)) % 4
        yield _ACGT[reads]


def frame_reads(chunks, read_len: int):
    """reads_chunks -> one framed sequence (every read followed by '\\n'), the layout kc_frame_fasta produces."""
    parts = []
    for r in chunks:
        b = np.empty((r.shape[0], read_len + 1), dtype=np.uint8)
        b[:, :read_len] = r
        b[:, read_len] = 10
        parts.append(b.reshape(-1))
    return np.concatenate(parts)


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


# Full-size BASELINE.json configurations with reference results in tests/golden/golden_big.json (written by
# tests/golden/make_golden_big.py in the build container, asserted by tests/test_gpu_big.py on the GPU box).
BIG_CONFIGS = {
    # three tiny ones (all word widths, with and without counts) that the CPU suite checks the oracle's digest against
    "tiny_k31z2": dict(config="tiny: 30x reads of a 20 kbp genome, k=31 -z 2", generator="reads_chunks(20_000, 30.0, 150, 0.01, 7)",
                       k=31, complements=True, min_frequency=2, one_line=True),
    "tiny_sparse_k31": dict(config="tiny: 3000 records of 31..34 bases (sparse switch, src/main.cpp:175)",
                            generator="short_records(3000, 31, 34, 17)", k=31, complements=True, min_frequency=1),
    "tiny_sparse_k9u": dict(config="tiny: 5000 records of 9..11 bases, k=9 -u (sparse switch, many overlaps)",
                            generator="short_records(5000, 9, 11, 19)", k=9, complements=False, min_frequency=1),
    "tiny_k63": dict(config="tiny: 200 kbp repeat-model genome, k=63", generator="human_like_genome(200_000, 11)",
                     k=63, complements=True, min_frequency=1),
    "tiny_k127u": dict(config="tiny: 200 kbp repeat-model genome, k=127 -u", generator="human_like_genome(200_000, 11)",
                       k=127, complements=False, min_frequency=1),
    "cfg1_50M": dict(config="configs[1] 50 x 1 Mbp uniform, k=31", generator="random_genome_records(50, 1_000_000, 12345)",
                     k=31, complements=True, min_frequency=1),
    "cfg2_k63u": dict(config="configs[2] 50 x 10 Mbp uniform, k=63 -u", generator="random_genome_records(50, 10_000_000, 31337)",
                      k=63, complements=False, min_frequency=1),
    "cfg2_k127u": dict(config="configs[2] 50 x 10 Mbp uniform, k=127 -u", generator="random_genome_records(50, 10_000_000, 31337)",
                       k=127, complements=False, min_frequency=1),
    "cfg3_reads_10M": dict(config="configs[3] at 1/10: 30x reads of a 10 Mbp genome, k=31 -z 2",
                           generator="reads_chunks(10_000_000, 30.0, 150, 0.01, 2024)", k=31, complements=True, min_frequency=2,
                           one_line=True),
    "cfg3_reads": dict(config="configs[3] 30x reads of a 100 Mbp genome, k=31 -z 2",
                       generator="reads_chunks(100_000_000, 30.0, 150, 0.01, 2024)", k=31, complements=True, min_frequency=2,
                       one_line=True),
    "cfg4_human_310M": dict(config="configs[4] at 1/10: 310 Mbp repeat-model genome, k=31", generator="human_like_genome(310_000_000, 3100)",
                            k=31, complements=True, min_frequency=1),
    "cfg4_human": dict(config="configs[4] 3.1 Gbp repeat-model genome, k=31", generator="human_like_genome(3_100_000_000, 3100)",
                       k=31, complements=True, min_frequency=1),
}


def big_config_input(name: str):
    """-> (seq, rec_off, rec_len) of BIG_CONFIGS[name], framed as kc_input wants it."""
    if name == "tiny_k31z2":
        seq = frame_reads(reads_chunks(20_000, 30.0, 150, 0.01, 7), 150)
        n = len(seq) // 151
        return seq, np.arange(n, dtype=np.uint64) * np.uint64(151), np.full(n, 150, dtype=np.uint64)
    if name == "tiny_sparse_k31":
        return frame_records(short_records(3000, 31, 34, 17))
    if name == "tiny_sparse_k9u":
        return frame_records(short_records(5000, 9, 11, 19))
    if name in ("tiny_k63", "tiny_k127u"):
        return frame_records(human_like_genome(200_000, 11))
    if name == "cfg1_50M":
        return frame_records(random_genome_records(50, 1_000_000, 12345))
    if name in ("cfg2_k63u", "cfg2_k127u"):
        return frame_records(random_genome_records(50, 10_000_000, 31337))
    if name in ("cfg3_reads", "cfg3_reads_10M"):
        glen = 100_000_000 if name == "cfg3_reads" else 10_000_000
        seq = frame_reads(reads_chunks(glen, 30.0, 150, 0.01, 2024), 150)
        n = len(seq) // 151
        return seq, np.arange(n, dtype=np.uint64) * np.uint64(151), np.full(n, 150, dtype=np.uint64)
    if name in ("cfg4_human", "cfg4_human_310M"):
        return frame_records(human_like_genome(3_100_000_000 if name == "cfg4_human" else 310_000_000, 3100))
    raise KeyError(name)

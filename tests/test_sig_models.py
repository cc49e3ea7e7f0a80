"""Python models of the index arithmetic the signature kernels rely on (kmercamel_b200/csrc/kmerset_sig.cuh).  The kernels themselves
are checked on the GPU (tests/test_gpu_sig.py: byte-identical superstrings with the exact construction); these models pin the two
bit-level identities they are built on, so that a change of either shows up without a GPU:

* kc_sig_scan_kernel cuts the runs of equal signature into pieces of <= 8 windows BIT-PARALLEL (three shifted ANDs) instead of
  walking the runs;
* kc_sig_resolve_kernel reads window t of a record and its reverse complement (reference src/parser.h:39-44: forward word, reverse
  complement, canonical = the smaller one) as static funnel shifts of the record's k + 7 bases F and of their reverse complement R.
"""
import random

M32 = 0xFFFFFFFF
M64 = (1 << 64) - 1


def pieces_by_walking(emr, neq):
    """the loop the scan used first: runs of valid windows of one signature, each cut every 8 windows from its start"""
    run_starts = emr & (neq | (~(emr << 1) & M32)) & M32
    stops = (run_starts | (~emr & M32)) & M32
    out = []
    rs = run_starts
    while rs:
        j = (rs & -rs).bit_length() - 1
        rs &= rs - 1
        rest = stops & ((M32 << (j + 1)) & M32) if j < 31 else 0
        ln = ((rest & -rest).bit_length() - 1 if rest else 32) - j
        while ln:
            l = min(ln, 8)
            out.append((j, l))
            j += l
            ln -= l
    return out


def pieces_bit_parallel(emr, neq):
    """kc_sig_scan_kernel: window p starts a piece iff it starts a run, or p - 8 starts a piece and windows p - 7 .. p continue its run"""
    run_starts = emr & (neq | (~(emr << 1) & M32)) & M32
    cont = emr & ~run_starts & M32
    c8 = cont & (cont << 1) & M32
    c8 &= (c8 << 2) & M32
    c8 &= (c8 << 4) & M32
    ps = run_starts
    q = (run_starts << 8) & c8 & M32
    ps |= q
    q = (q << 8) & c8 & M32
    ps |= q
    q = (q << 8) & c8 & M32
    ps |= q
    bound = (ps | (~emr & M32)) & M32
    out = []
    while ps:
        j = (ps & -ps).bit_length() - 1
        ps &= ps - 1
        rest = bound & ((0xFFFFFFFE << j) & M32)
        out.append((j, ((rest & -rest).bit_length() - 1 if rest else 32) - j))
    return out


def test_piece_cutting_bit_parallel_equals_walking():
    rnd = random.Random(1)
    for it in range(60000):
        mode = it % 4
        emr = rnd.getrandbits(32) if mode == 0 else (M32 if mode == 1 else rnd.getrandbits(32) | rnd.getrandbits(32) | rnd.getrandbits(32))
        neq = (rnd.getrandbits(32) & rnd.getrandbits(32) & rnd.getrandbits(32) if mode != 3 else 0) | 1
        assert pieces_by_walking(emr, neq) == pieces_bit_parallel(emr, neq), (hex(emr), hex(neq))
    for emr, neq in [(M32, 1), (0, 1), (1, 1), (0x80000000, 1), (M32, M32), (0xFFFF0000, 1), (0x0001FFFE, 1)]:
        got = pieces_bit_parallel(emr, neq)
        assert got == pieces_by_walking(emr, neq)
        assert all(1 <= l <= 8 for _, l in got) and sum(l for _, l in got) == bin(emr).count("1")


def rs64(w):  # kc_reverse_symbols64_brev: bit reversal, then the two bits of every symbol swapped back
    b = int("{:064b}".format(w)[::-1], 2)
    return ((b >> 1) & 0x5555555555555555) | ((b & 0x5555555555555555) << 1)


def revcomp(x, k):
    r = 0
    for i in range(k):
        r = (r << 2) | (3 - ((x >> (2 * i)) & 3))
    return r


def windows_by_funnel(bases, k, L):
    """the spread step of kc_sig_resolve_kernel for L limbs: -> [(forward window t, its reverse complement)] for t = 0..7"""
    f = 0
    for b in bases[:k]:
        f = (f << 2) | b
    ms = 0
    for b in bases[k:]:
        ms = (ms << 2) | b
    ms <<= 64 - 14                                    # the 7 bases behind the first window, left-aligned in 64 bits
    fl = [(f >> (64 * i)) & M64 for i in range(L)]
    F = [0] * (L + 1)
    F[0] = ((fl[0] << 14) | (ms >> 50)) & M64
    for i in range(1, L):
        F[i] = ((fl[i] << 14) | (fl[i - 1] >> 50)) & M64
    F[L] = fl[L - 1] >> 50
    Rw = [(~rs64(F[L - i])) & M64 for i in range(L + 1)]
    Rv = sum(Rw[i] << (64 * i) for i in range(L + 1)) >> (64 * (L + 1) - 2 * (k + 7))
    R = [(Rv >> (64 * i)) & M64 for i in range(L + 1)]
    kmask = (1 << (2 * k)) - 1
    out = []
    for t in range(8):
        sf, sr = 14 - 2 * t, 2 * t
        ft = sum((((F[j] >> sf) | ((F[j + 1] << (64 - sf)) if sf else 0)) & M64) << (64 * j) for j in range(L)) & kmask
        rt = sum((((R[j] >> sr) | ((R[j + 1] << (64 - sr)) if sr else 0)) & M64) << (64 * j) for j in range(L)) & kmask
        out.append((ft, rt))
    return out


def test_windows_by_static_funnel_shifts():
    rnd = random.Random(7)
    for L, ks in ((1, range(26, 32)), (2, (32, 33, 40, 47, 62, 63)), (4, (64, 65, 95, 96, 126, 127))):
        for k in ks:
            for _ in range(40):
                bases = [rnd.randrange(4) for _ in range(k + 7)]
                for t, (ft, rt) in enumerate(windows_by_funnel(bases, k, L)):
                    want = 0
                    for b in bases[t:t + k]:
                        want = (want << 2) | b
                    assert ft == want and rt == revcomp(want, k), (L, k, t)


def test_decode_validity_by_rebuilding_the_letter():
    """kc_pack4 (csrc/stage1.cuh): a byte is a nucleotide iff its case-folded value equals the letter its 2-bit code stands for"""
    def valid4(w):
        c = ((w >> 1) ^ (w >> 2)) & 0x03030303
        c0, c1 = c & 0x01010101, (c >> 1) & 0x01010101
        expect = (0x41414141 + c0 * 2 + c1 * 6 + (c0 & c1) * 11) & M32
        d = (w & 0xDFDFDFDF) ^ expect
        z = ~(((d & 0x7F7F7F7F) + 0x7F7F7F7F) | d) & 0x80808080 & M32
        return ((((z >> 7) * 0x08040201) & M32) >> 24) & 0xF
    def want4(w):
        return sum(1 << (3 - i) for i in range(4) if chr((w >> (8 * i)) & 0xFF) in "ACGTacgt")
    for b in range(256):
        for pos in range(4):
            for other in (0x41, 0x54, 0x00, 0xFF, 0x4E, 0x67):
                w = sum((b if i == pos else other) << (8 * i) for i in range(4))
                assert valid4(w) == want4(w), hex(w)
    rnd = random.Random(3)
    for _ in range(20000):
        w = rnd.getrandbits(32)
        assert valid4(w) == want4(w), hex(w)

// KWord<L>: fixed-width unsigned integer of L little-endian 64-bit limbs, the device-side k-mer word.
//
// Replaces the reference's kmer64_t / kmer128_t / uint256_t (src/kmers.h:11-13, src/uint256_t/) with one
// register-resident type: L = 1 (k < 32), 2 (k < 64), 4 (k < 128), the widths chosen at src/main.cpp:309-315.
// Layout of a k-mer is the reference's: base i (0 = leftmost) at bits 2(k-1-i)+1..2(k-1-i), A=0 C=1 G=2 T=3,
// right-aligned, so unsigned integer order is lexicographic order.  Wider L (3, 5) is used for sort tuples.
//
// All limb indices are compile-time constants after unrolling, so words never spill to local memory.
#pragma once
#include "kc_common.cuh"

template <int L> struct KWord {
    u64 w[L];

    KC_HD static KWord zero() {
        KWord r;
#pragma unroll
        for (int i = 0; i < L; ++i) r.w[i] = 0;
        return r;
    }
    KC_HD static KWord ones() {
        KWord r;
#pragma unroll
        for (int i = 0; i < L; ++i) r.w[i] = ~0ULL;
        return r;
    }
    KC_HD static KWord from_u64(u64 v) {
        KWord r = zero();
        r.w[0] = v;
        return r;
    }
    // (1 << bits) - 1 for 0 <= bits <= 64 L
    KC_HD static KWord low_mask(int bits) {
        KWord r;
#pragma unroll
        for (int i = 0; i < L; ++i) {
            int rem = bits - 64 * i;
            r.w[i] = rem >= 64 ? ~0ULL : (rem > 0 ? ((1ULL << rem) - 1) : 0ULL);
        }
        return r;
    }
    KC_HD bool operator==(const KWord &o) const {
        bool e = true;
#pragma unroll
        for (int i = 0; i < L; ++i) e = e && (w[i] == o.w[i]);
        return e;
    }
    KC_HD bool operator!=(const KWord &o) const { return !(*this == o); }
    KC_HD bool operator<(const KWord &o) const {
        bool lt = false, decided = false;
#pragma unroll
        for (int i = L - 1; i >= 0; --i) {
            if (!decided && w[i] != o.w[i]) {
                lt = w[i] < o.w[i];
                decided = true;
            }
        }
        return lt;
    }
    KC_HD bool operator<=(const KWord &o) const { return !(o < *this); }
    KC_HD KWord operator|(const KWord &o) const {
        KWord r;
#pragma unroll
        for (int i = 0; i < L; ++i) r.w[i] = w[i] | o.w[i];
        return r;
    }
    KC_HD KWord operator&(const KWord &o) const {
        KWord r;
#pragma unroll
        for (int i = 0; i < L; ++i) r.w[i] = w[i] & o.w[i];
        return r;
    }
    KC_HD KWord operator~() const {
        KWord r;
#pragma unroll
        for (int i = 0; i < L; ++i) r.w[i] = ~w[i];
        return r;
    }
    // Logical shift right by 0 <= bits (>= 64 L gives 0).
    KC_HD KWord shr(int bits) const {
        const int limb = bits >> 6, off = bits & 63;
        KWord t = *this;
        if (L > 1 && limb) {
            t = zero();
#pragma unroll
            for (int s = 1; s < L; ++s)
                if (limb == s) {
#pragma unroll
                    for (int i = 0; i + s < L; ++i) t.w[i] = w[i + s];
                }
        } else if (L == 1 && limb) {
            t = zero();
        }
        if (off) {
#pragma unroll
            for (int i = 0; i < L; ++i) t.w[i] = (t.w[i] >> off) | (i + 1 < L ? (t.w[i + 1] << (64 - off)) : 0ULL);
        }
        return t;
    }
    // Logical shift left by 0 <= bits.
    KC_HD KWord shl(int bits) const {
        const int limb = bits >> 6, off = bits & 63;
        KWord t = *this;
        if (L > 1 && limb) {
            t = zero();
#pragma unroll
            for (int s = 1; s < L; ++s)
                if (limb == s) {
#pragma unroll
                    for (int i = L - 1; i - s >= 0; --i) t.w[i] = w[i - s];
                }
        } else if (L == 1 && limb) {
            t = zero();
        }
        if (off) {
#pragma unroll
            for (int i = L - 1; i >= 0; --i) t.w[i] = (t.w[i] << off) | (i > 0 ? (t.w[i - 1] >> (64 - off)) : 0ULL);
        }
        return t;
    }
    // Bits [pos, pos+nbits) as an integer, nbits <= 32.
    KC_HD u32 bits_at(int pos, int nbits) const { return (u32) (shr(pos).w[0] & ((1ULL << nbits) - 1)); }
    // Same when [pos, pos+nbits) does not straddle more than two limbs and pos is dynamic: cheaper than shr().
    KC_HD u32 digit(int pos, int nbits) const {
        const int limb = pos >> 6, off = pos & 63;
        u64 lo = 0, hi = 0;
#pragma unroll
        for (int i = 0; i < L; ++i) {
            if (i == limb) lo = w[i];
            if (i == limb + 1) hi = w[i];
        }
        u64 v = lo >> off;
        if (off) v |= hi << (64 - off);
        return (u32) (v & ((1ULL << nbits) - 1));
    }
    // Same when the digit lies inside the TOP limb (pos >= 64 (L - 1), pos + nbits <= 64 L): the MSD digits of the fixed-slot
    // partition levels.  One 64-bit shift + mask instead of the limb selection above (ncu r01h: digit() was 15-18 % of the
    // instructions of the two scatter kernels, three calls per item).
    KC_HD u32 digit_top(int pos, int nbits) const { return (u32) (w[L - 1] >> (pos - 64 * (L - 1))) & ((1u << nbits) - 1u); }
};

// ---- k-mer arithmetic (reference src/kmers.h) -------------------------------------------------------

// src/kmers.h:35-38 BitPrefix: the first d bases of a k-mer.
template <int L> KC_HD KWord<L> kmer_prefix(const KWord<L> &x, int k, int d) { return x.shr(2 * (k - d)); }
// src/kmers.h:41-44 BitSuffix: the last d bases.
template <int L> KC_HD KWord<L> kmer_suffix(const KWord<L> &x, int d) { return x & KWord<L>::low_mask(2 * d); }

// Reverse the 32 two-bit symbols of one limb (the swap network of src/kmers.h:61-70 on 64 bits).
KC_HD u64 kc_reverse_symbols64(u64 w) {
    w = ((w >> 2) & 0x3333333333333333ULL) | ((w & 0x3333333333333333ULL) << 2);
    w = ((w >> 4) & 0x0F0F0F0F0F0F0F0FULL) | ((w & 0x0F0F0F0F0F0F0F0FULL) << 4);
    w = ((w >> 8) & 0x00FF00FF00FF00FFULL) | ((w & 0x00FF00FF00FF00FFULL) << 8);
    w = ((w >> 16) & 0x0000FFFF0000FFFFULL) | ((w & 0x0000FFFF0000FFFFULL) << 16);
    return (w >> 32) | (w << 32);
}

// src/kmers.h:92-95 ReverseComplement of a k-mer (also used for d-mers with k := d).
template <int L> KC_HD KWord<L> kmer_reverse_complement(const KWord<L> &x, int k) {
    KWord<L> r;
#pragma unroll
    for (int i = 0; i < L; ++i) r.w[i] = ~kc_reverse_symbols64(x.w[L - 1 - i]);
    return r.shr(64 * L - 2 * k);
}

// src/kmers.h:99-102 AtIndex: the symbol (0..3) at position index (0 = leftmost).
template <int L> KC_HD u32 kmer_symbol(const KWord<L> &x, int k, int index) { return x.digit(2 * (k - 1 - index), 2); }

// src/kmers.h:15-32 nucleotideToInt without the table: A/a C/c G/g T/t -> 0..3, anything else 4.
KC_HD u32 kc_nucleotide_code(u32 c) {
    u32 u = c & 0xDFu;  // fold case
    u32 code = 4;
    if (u == 'A') code = 0;
    if (u == 'C') code = 1;
    if (u == 'G') code = 2;
    if (u == 'T') code = 3;
    return code;
}


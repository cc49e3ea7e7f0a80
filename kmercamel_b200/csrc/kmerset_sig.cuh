// K-mer set construction by SIGNATURE buckets (super-k-mers) for the FLAGS regime: the same result as kmerset_fast.cuh /
// kmerset.cuh — bit p of the flag array set iff the k-mer of the window ending at p occurs there for the first time
// (reference src/parser.h:22-141 AddKMers / ReadKMers + khash) — with ~5 x less HBM traffic and two kernels instead of three.
//
// The partition passes of kmerset_fast.cuh move every k-mer occurrence as a 12-byte (key, position) item: 1 + 12 (level 0)
// + 12 + 12 (level 1) + 12 (leaf resolve) = 49 bytes per k-mer, and their cost is the shared-memory ranking of those items.
// Here the unit that travels is a RECORD: a run of consecutive windows of one 32-position strip whose k-mers share a
// signature, 8 bytes for ~7 windows.  The k-mers themselves are re-created from the sequence (which stays in HBM anyway: the
// emission reads it again) inside the CTA that resolves the bucket.
//
//   signature of a k-mer  = the smallest hash among the canonical M-mers at its W = 16 CENTRAL M-mer positions
//                           (M = 16 / 15 for k >= 30, 12 / 11 for 26 <= k < 30; chosen so that k - M - (W - 1) is even:
//                           the central range is then mirror-symmetric, a k-mer and its reverse complement see the same
//                           set of canonical M-mers, and every occurrence of a canonical k-mer gets the same signature);
//   bucket                = floor(n_buckets * (1 - (1 - s)^14)), s = signature / 2^32.  The minimum of W uniform hashes has
//                           distribution 1 - (1 - s)^W; exponent W would spread the WINDOWS evenly over the buckets, but a large
//                           signature changes hands after ~2 windows where a small one lasts for ~16, so the last buckets would
//                           hold four times the RECORDS of the first ones.  Exponent 14 gives the first buckets 14 % more windows
//                           and the last ones fewer, which evens out the records (simulation: windows max / mean 1.43, records
//                           max / mean 1.6 over 3.5 k buckets instead of 1.3 and 3.0);
//   kc_sig_scan_kernel    sequence -> 2-bit codes (as level 0 of the other constructions) -> 47 M-mer hashes per strip by static
//                           funnel shifts of the strip's 128-bit code window and of its reverse complement -> sliding minimum by
//                           doubling (2, 4, 8, 16) in registers -> bucket per window -> one record per run: ONE 64-bit
//                           atomicAdd on the bucket's cursor reserves the record slot (low word) and the item range (high word),
//                           one 8-byte store writes {position, length, first item}.  Also writes the valid-window bits
//                           ("clear the losers" flags, kmerset_fast.cuh).
//   kc_sig_resolve_kernel persistent CTAs, one bucket at a time (<= CAP windows): a thread per record decodes the record's
//                           strip (+ L strips to the left) into shared memory; a thread per window funnel-shifts its k-mer out of
//                           those words, takes the canonical form and resolves it in the shared-memory tables of
//                           kc_ksf_resolve2_kernel (plain-store table T1, one barrier, CAS fallback T2); every folded duplicate
//                           clears the bit of the larger position.
//
// Algorithmic HBM bytes per k-mer (L = 1): 1 (sequence) + ~1.2 (record written) + ~1.2 (record read) + ~9.6 (the two 32-byte
// sectors a record's strip is re-read as, ~6.7 windows per record) = ~13 instead of 49.
//
// A bucket that exceeds its capacity (records or windows: heavily repeated sequence) sets the status word; the caller
// discards the flags and uses the other constructions, exactly like an overflow in kmerset_fast.cuh.
#pragma once
#include "kmerset_fast.cuh"

struct SigTuning {
    bool enabled = true;
    uint32_t load_pct = 0;             // mean windows per bucket, in percent of the bucket capacity; 0 = 50 / 44 / 39 for L = 1 / 2 / 4
                                       // (the first buckets get 14 % more, sigma ~ 4.2 sqrt(mean) because windows arrive in runs: the
                                       // mean + 5.5 sigma that the largest of 10^6 buckets reaches stays below the capacity)
    uint64_t min_items = 1u << 16;     // smaller inputs: nothing to gain
};

template <int L> struct SigCfg {
    static constexpr u32 CAP = L == 1 ? 4096 : (L == 2 ? 2048 : 1024);  // windows (items) per bucket
    static constexpr u32 REC_CAP = CAP / 4;                              // records per bucket
    static constexpr int THREADS = L == 1 ? 512 : 256;
    static constexpr int NW = L + 1;                                     // 32-base words a window can reach into
    static constexpr int MIN_CTAS = L == 1 ? 2 : 3;                      // resident CTAs per SM the shared memory allows
};

static const int KC_SIG_W = 16;

struct SigPlan {
    bool ok = false;
    int m = 0;          // M-mer length
    int a = 0;          // the central range of a window ending at e: M-mers ending at e - a - (W - 1) .. e - a
    u32 n_buckets = 0;
};

inline SigPlan kc_sig_plan(u64 n_bytes, int k, u32 cap, const SigTuning &t) {
    SigPlan p;
    if (!t.enabled || n_bytes < t.min_items || n_bytes >= 0xFFFFFFFFULL) return p;
    static const int ms[4] = {16, 15, 12, 11};
    for (int i = 0; i < 4 && !p.m; ++i) {
        const int d = k - ms[i] - (KC_SIG_W - 1);
        if (d >= 0 && d % 2 == 0) {
            p.m = ms[i];
            p.a = d / 2;
        }
    }
    if (!p.m) return p;
    const u64 pct = t.load_pct ? t.load_pct : (cap >= 4096 ? 50 : (cap >= 2048 ? 44 : 39));
    const u64 mean = (u64) cap * pct / 100;
    const u64 nb = kc_div_up(n_bytes, mean ? mean : 1);
    if (nb >= (1u << 26)) return p;  // a record keeps its bucket's item base in 26 bits... and the scan its bucket in 27
    p.n_buckets = (u32) (nb < 64 ? 64 : nb);
    p.ok = true;
    return p;
}

#ifdef __CUDACC__

// 4 ASCII bytes (byte 0 = first base) -> 8 bits of codes, first base in the top 2 bits (kc_pack4 without the validity bits).
KC_D u32 kc_codes4(u32 w) {
    const u32 c = ((w >> 1) ^ (w >> 2)) & 0x03030303u;
    return (c * 0x40100401u) >> 24;
}

// Reverse complement of 16 bases held in one 32-bit word.
KC_D u32 kc_revcomp16(u32 w) {
    u32 b = __brev(w);  // symbols reversed, the two bits of every symbol swapped
    b = ((b >> 1) & 0x55555555u) | ((b & 0x55555555u) << 1);
    return ~b;
}

// Bits [SH, SH + 31] of the 128-bit value x[0] : x[1] : x[2] : x[3] (x[0] most significant), SH static.
template <int SH> KC_D u32 kc_bits128(const u32 (&x)[4]) {
    constexpr int wr = SH >> 5, off = SH & 31;  // word index from the right
    const u32 lo = x[3 - wr];
    if constexpr (off == 0) {
        return lo;
    } else {
        const u32 hi = wr + 1 <= 3 ? x[3 - (wr + 1 <= 3 ? wr + 1 : 3)] : 0u;
        return __funnelshift_r(lo, hi, off);
    }
}

template <int M, int I> struct SigHashLoop {
    KC_D static void run(const u32 (&x)[4], const u32 (&y)[4], u32 *H) {
        constexpr u32 mask = M == 16 ? 0xFFFFFFFFu : ((1u << (2 * M)) - 1u);
        const u32 f = kc_bits128<128 - 2 * (I + M)>(x) & mask;  // bases I .. I + M - 1 of the code window
        const u32 r = kc_bits128<2 * I>(y) & mask;               // their reverse complement
        u32 h = (f < r ? f : r) * 0x9E3779B1u;
        h ^= h >> 15;
        h *= 0x85EBCA6Bu;
        H[I] = h;
        SigHashLoop<M, I + 1>::run(x, y, H);
    }
};
template <int M> struct SigHashLoop<M, 32 + KC_SIG_W - 1> {
    KC_D static void run(const u32 (&)[4], const u32 (&)[4], u32 *) {}
};

// A record: bits 0..31 END position of its first window, 32..36 windows - 1, 38..63 index of its first item inside the bucket.
KC_HD u64 kc_sig_record(u32 pos, u32 len, u32 item_base) { return (u64) pos | ((u64) (len - 1) << 32) | ((u64) item_base << 38); }

template <int M>
__global__ void __launch_bounds__(256) kc_sig_scan_kernel(const u8 *__restrict__ seq, u64 n_bytes, int k, int a, u32 n_buckets, u32 rec_cap, u32 item_cap,
                                                          kc_ull *cursor, u64 *__restrict__ recs, u32 *__restrict__ flags, u32 n_flag_words, u32 tile0,
                                                          kc_ull *m_cell, u32 *status, const u32 *__restrict__ win_mask) {
    constexpr int T = 256;
    constexpr int NH = 32 + KC_SIG_W - 1;  // M-mer hashes a strip needs
    __shared__ u64 pk[KC_EX_HALO + T];
    __shared__ u32 vm[KC_EX_HALO + T];
    __shared__ u32 starts[KC_EX_STRIP * T];  // (bucket << 5 | first window) of the thread's c-th record at [c * T + thread]
    const i64 block_pos0 = (i64) (tile0 + blockIdx.x) * (T * KC_EX_STRIP);
    kc_tile_load<T>(seq, n_bytes, block_pos0, pk, vm);
    __syncthreads();
    const int widx = KC_EX_HALO + threadIdx.x;
    const u32 em = kc_strip_emit_mask(vm, widx, k) & kc_strip_window_filter(win_mask, block_pos0);
    {   // clear-the-losers flags: every window starts as "first occurrence" (bit p & 31 of word p >> 5 = window END p)
        const u64 w = (u64) (block_pos0 >> 5) + threadIdx.x;
        if (em && w < n_flag_words) flags[w] = __brev(em);
    }
    {
        u32 c = __popc(em);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xFFFFFFFFu, c, o);
        if ((threadIdx.x & 31) == 0 && c) atomicAdd(m_cell, (kc_ull) c);
    }
    if (!em) return;
    // the 64 bases that end with the strip's last M-mer: base t of the code window = strip base t - back
    const int back = a + (KC_SIG_W - 2) + M;
    const int g0 = widx * 32 - back;
    const int wi0 = g0 >> 5, sb = 2 * (g0 & 31);
    const u64 p0 = pk[wi0], p1 = pk[wi0 + 1], p2 = wi0 + 2 <= widx ? pk[wi0 + 2] : 0ULL;
    const u64 X0 = sb ? (p0 << sb) | (p1 >> (64 - sb)) : p0;
    const u64 X1 = sb ? (p1 << sb) | (p2 >> (64 - sb)) : p1;
    const u32 x[4] = {(u32) (X0 >> 32), (u32) X0, (u32) (X1 >> 32), (u32) X1};
    const u32 y[4] = {kc_revcomp16(x[3]), kc_revcomp16(x[2]), kc_revcomp16(x[1]), kc_revcomp16(x[0])};
    u32 H[NH];
    SigHashLoop<M, 0>::run(x, y, H);
    // sliding minimum over W = 16 hashes by doubling; H[j] = signature of the window ending at strip base j
#pragma unroll
    for (int i = 0; i < NH - 1; ++i) H[i] = min(H[i], H[i + 1]);
#pragma unroll
    for (int i = 0; i < NH - 3; ++i) H[i] = min(H[i], H[i + 2]);
#pragma unroll
    for (int i = 0; i < NH - 7; ++i) H[i] = min(H[i], H[i + 4]);
#pragma unroll
    for (int i = 0; i < NH - 15; ++i) H[i] = min(H[i], H[i + 8]);
    const u32 emr = __brev(em);  // bit j = window ending at strip base j
    u32 n_rec = 0, prev_b = 0;
#pragma unroll
    for (int j = 0; j < KC_EX_STRIP; ++j) {
        const u32 u1 = ~H[j];
        const u32 u2 = __umulhi(u1, u1), u4 = __umulhi(u2, u2), u8 = __umulhi(u4, u4);
        const u32 u = __umulhi(__umulhi(u8, u4), u2);  // (1 - s)^14
        const u32 b = __umulhi(~u, n_buckets);
        const bool valid = (emr >> j) & 1u;
        const bool first = valid && (j == 0 || !((emr >> (j - 1)) & 1u) || b != prev_b);
        if (first) {
            starts[n_rec * T + threadIdx.x] = (b << 5) | (u32) j;
            ++n_rec;
        }
        prev_b = b;
    }
    // stops: a record ends where the next one starts or where the valid windows end
    u32 bnd = 0;
    for (u32 c = 0; c < n_rec; ++c) bnd |= 1u << (starts[c * T + threadIdx.x] & 31u);
    const u32 stops = bnd | ~emr;
    const u32 pos0 = (u32) block_pos0 + threadIdx.x * KC_EX_STRIP;
    bool over = false;
    for (u32 c = 0; c < n_rec; ++c) {
        const u32 e = starts[c * T + threadIdx.x];
        const u32 j = e & 31u, b = e >> 5;
        const u32 rest = j < 31 ? stops & (0xFFFFFFFFu << (j + 1)) : 0u;
        const u32 len = (rest ? (u32) __ffs(rest) - 1u : 32u) - j;
        const kc_ull old = atomicAdd(&cursor[b], ((kc_ull) len << 32) | 1ULL);
        const u32 slot = (u32) old, ib = (u32) (old >> 32);
        if (slot < rec_cap && ib + len <= item_cap) recs[(u64) b * rec_cap + slot] = kc_sig_record(pos0 + j, len, ib);
        else over = true;
    }
    if (over) status[0] = 1;
}

template <int L, bool MULTI>
__global__ void __launch_bounds__(SigCfg<L>::THREADS, SigCfg<L>::MIN_CTAS) kc_sig_resolve_kernel(const u8 *__restrict__ seq, u64 n_bytes, int k, int complements,
                                                                             const kc_ull *__restrict__ cursor, const u64 *__restrict__ recs, u32 n_buckets,
                                                                             KsfFlagPeers fl, kc_ull *n_unique, u32 *status) {
    typedef SigCfg<L> Cfg;
    constexpr u32 CAP = Cfg::CAP, REC_CAP = Cfg::REC_CAP;
    constexpr int T = Cfg::THREADS, NW = Cfg::NW;
    constexpr int NI = (int) CAP / T;
    constexpr u32 T1N = 2 * CAP, T2N = CAP;
    constexpr int T1_BITS = CAP == 4096 ? 13 : (CAP == 2048 ? 12 : 11);
    static_assert((1u << T1_BITS) == T1N, "T1 size");
    static_assert(NI % 4 == 0 && NI * T == (int) CAP, "a thread clears NI slots of T2 with 16-byte stores");
    extern __shared__ __align__(16) unsigned char kc_smem_raw[];
    KWord<L> *sk = reinterpret_cast<KWord<L> *>(kc_smem_raw);  // [CAP] canonical k-mer of item i
    u64 *swords = reinterpret_cast<u64 *>(sk + CAP);            // [NW][REC_CAP] code words of record r: word NW - 1 = its own strip
    u32 *sp = reinterpret_cast<u32 *>(swords + NW * REC_CAP);   // [CAP] window END position of item i (the smallest one of its key, once folded)
    u32 *T2 = sp + CAP;                                         // [T2N]
    u32 *recinfo = T2 + T2N;                                    // [REC_CAP] position of the record's first window - its first item
    u16 *T1 = reinterpret_cast<u16 *>(recinfo + REC_CAP);       // [T1N]
    u16 *map = T1 + T1N;                                        // [CAP] item -> record
    const KWord<L> kmask = KWord<L>::low_mask(2 * k);
    u32 kept = 0;
    for (u32 b = blockIdx.x; b < n_buckets; b += gridDim.x) {
        const u64 cur = cursor[b];
        const u32 n_rec = (u32) cur, n_items = (u32) (cur >> 32);
        if (n_items == 0) continue;
        if (n_rec > REC_CAP || n_items > CAP) {  // the scan has set the status word already
            if (threadIdx.x == 0) status[0] = 1;
            continue;
        }
        // P0: a thread per record
        for (u32 r = threadIdx.x; r < n_rec; r += T) {
            const u64 rec = recs[(u64) b * REC_CAP + r];
            const u32 pos = (u32) rec, len = ((u32) (rec >> 32) & 31u) + 1u, ib = (u32) (rec >> 38);
            const i64 s0 = (i64) (pos & ~31u);
#pragma unroll
            for (int wi = 0; wi < NW; ++wi) {
                const i64 p = s0 - 32 * (NW - 1 - wi);
                u64 codes = 0;
                if (p >= 0 && (u64) p + 32 <= n_bytes) {
                    const uint4 *src = reinterpret_cast<const uint4 *>(seq + p);
                    const uint4 a4 = __ldg(src), b4 = __ldg(src + 1);
                    const u32 w[8] = {a4.x, a4.y, a4.z, a4.w, b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                    for (int j = 0; j < 8; ++j) codes = (codes << 8) | kc_codes4(w[j]);
                } else if (p + 32 > 0 && (u64) (p < 0 ? 0 : p) < n_bytes) {
                    for (int j = 0; j < 32; ++j) {
                        const i64 q = p + j;
                        u32 code = 0;
                        if (q >= 0 && (u64) q < n_bytes) code = kc_nucleotide_code(seq[q]) & 3u;
                        codes = (codes << 2) | code;
                    }
                }
                swords[wi * REC_CAP + r] = codes;
            }
            recinfo[r] = pos - ib;
            for (u32 t = 0; t < len; ++t) map[ib + t] = (u16) r;
        }
        __syncthreads();
        // P1: a thread per window: k-mer out of the record's words, canonical, into the tables
        {
            uint4 *t2v = reinterpret_cast<uint4 *>(T2) + threadIdx.x * (NI / 4);
#pragma unroll
            for (int q = 0; q < NI / 4; ++q) t2v[q] = make_uint4(KC_NONE, KC_NONE, KC_NONE, KC_NONE);
        }
        KWord<L> key[NI];
        u32 h1[NI], h2[NI];
#pragma unroll
        for (int m = 0; m < NI; ++m) {
            const u32 i = threadIdx.x + (u32) m * T;
            if (i < n_items) {
                const u32 r = map[i];
                const u32 p = recinfo[r] + i;
                const int sh = 2 * (31 - (int) (p & 31u));
                KWord<L> f;
#pragma unroll
                for (int l = 0; l < L; ++l) {
                    const u64 lo = swords[(NW - 1 - l) * REC_CAP + r], hi = swords[(NW - 2 - l) * REC_CAP + r];
                    f.w[l] = sh ? (lo >> sh) | (hi << (64 - sh)) : lo;
                }
                f = f & kmask;
                if (complements) {
                    const KWord<L> rc = kmer_reverse_complement(f, k);
                    if (rc < f) f = rc;
                }
                key[m] = f;
                sk[i] = f;
                sp[i] = p;
                u64 h = 0;
#pragma unroll
                for (int w = 0; w < L; ++w) h = (h ^ f.w[w]) * 0xD6E8FEB86659FD93ULL;
                h1[m] = (u32) (h >> (64 - T1_BITS));
                h2[m] = (u32) (h >> (64 - 2 * T1_BITS)) & (T2N - 1);
            }
        }
#pragma unroll
        for (int m = 0; m < NI; ++m) {
            const u32 i = threadIdx.x + (u32) m * T;
            if (i < n_items) T1[h1[m]] = (u16) i;
        }
        __syncthreads();
        // P2: the winner of a T1 slot represents its key; equal key = duplicate (fold the positions, clear the larger one's bit);
        //     different key -> T2 with CAS + linear probing
        u32 o[NI];
#pragma unroll
        for (int m = 0; m < NI; ++m) {
            const u32 i = threadIdx.x + (u32) m * T;
            o[m] = i < n_items ? (u32) T1[h1[m]] : i;
        }
#pragma unroll
        for (int m = 0; m < NI; ++m) {
            const u32 i = threadIdx.x + (u32) m * T;
            if (i >= n_items) continue;
            if (o[m] == i) {
                ++kept;
            } else if (sk[o[m]] == key[m]) {
                const u32 mine = sp[i];
                const u32 was = atomicMin(&sp[o[m]], mine);
                kc_flag_clear_all<MULTI>(fl, was > mine ? was : mine);
            } else {
                u32 s = h2[m];
                while (true) {
                    const u32 old = atomicCAS(&T2[s], KC_NONE, i);
                    if (old == KC_NONE) {
                        ++kept;
                        break;
                    }
                    if (sk[old] == key[m]) {
                        const u32 mine = sp[i];
                        const u32 was = atomicMin(&sp[old], mine);
                        kc_flag_clear_all<MULTI>(fl, was > mine ? was : mine);
                        break;
                    }
                    s = (s + 1) & (T2N - 1);
                }
            }
        }
        // no barrier here: P0 of the next bucket writes swords / recinfo / map only, which P2 does not read, and its barrier
        // keeps P1 (sk, sp, T1, T2) behind every thread's P2
    }
#pragma unroll
    for (int o2 = 16; o2 > 0; o2 >>= 1) kept += __shfl_down_sync(0xFFFFFFFFu, kept, o2);
    if ((threadIdx.x & 31) == 0 && kept) atomicAdd(n_unique, (kc_ull) kept);
}

template <int L> struct SigKernels {
    typedef SigCfg<L> Cfg;
    static int smem_r() {
        return (int) (Cfg::CAP * (sizeof(KWord<L>) + 4 + 4 + 2 * 2 + 2) + Cfg::REC_CAP * (Cfg::NW * 8 + 4));
    }
    struct Dev {
        int n_sm = 0, occ = 0;
    };
    static const Dev &prepare() {
        static KcDevOnce once;
        static Dev dev[KC_MAX_DEVICES];
        const int d = once.run([&](int dv) {
            KC_CUDA(cudaFuncSetAttribute(kc_sig_resolve_kernel<L, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_r()));
            KC_CUDA(cudaFuncSetAttribute(kc_sig_resolve_kernel<L, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_r()));
            dev[dv].n_sm = kc_sm_count(dv);
            KC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&dev[dv].occ, kc_sig_resolve_kernel<L, false>, Cfg::THREADS, smem_r()));
        });
        return dev[d];
    }
};

inline void kc_sig_scan_launch(int m, u32 blocks, cudaStream_t st, const u8 *seq, u64 n_bytes, int k, int a, u32 n_buckets, u32 rec_cap, u32 item_cap,
                               kc_ull *cursor, u64 *recs, u32 *flags, u32 n_flag_words, u32 tile0, kc_ull *m_cell, u32 *status, const u32 *win_mask) {
    switch (m) {
    case 16: kc_sig_scan_kernel<16><<<blocks, 256, 0, st>>>(seq, n_bytes, k, a, n_buckets, rec_cap, item_cap, cursor, recs, flags, n_flag_words, tile0, m_cell, status, win_mask); break;
    case 15: kc_sig_scan_kernel<15><<<blocks, 256, 0, st>>>(seq, n_bytes, k, a, n_buckets, rec_cap, item_cap, cursor, recs, flags, n_flag_words, tile0, m_cell, status, win_mask); break;
    case 12: kc_sig_scan_kernel<12><<<blocks, 256, 0, st>>>(seq, n_bytes, k, a, n_buckets, rec_cap, item_cap, cursor, recs, flags, n_flag_words, tile0, m_cell, status, win_mask); break;
    default: kc_sig_scan_kernel<11><<<blocks, 256, 0, st>>>(seq, n_bytes, k, a, n_buckets, rec_cap, item_cap, cursor, recs, flags, n_flag_words, tile0, m_cell, status, win_mask); break;
    }
    KC_CUDA(cudaGetLastError());
}

// Launches the construction on ex.stream and returns without synchronising; same contract as kc_kmerset_build_fast:
//   cells[0] += distinct k-mers, cells[2] = M (k-mer windows), low word of cells[3] = overflow status; flags arrive zeroed.
// Returns false (nothing launched) when the plan does not apply.
template <int L>
bool kc_kmerset_build_sig(CudaExec &ex, const u8 *seq, u64 n_bytes, int k, bool complements, u32 *flags, u64 *cells, const SigTuning &tune,
                          InputChunks *chunks = nullptr, const u32 *win_mask = nullptr, SigPlan *plan_out = nullptr) {
    typedef SigCfg<L> Cfg;
    typedef SigKernels<L> KK;
    const SigPlan pl = kc_sig_plan(n_bytes, k, Cfg::CAP, tune);
    if (plan_out) *plan_out = pl;
    if (!pl.ok) return false;
    const typename KK::Dev &dv = KK::prepare();
    cudaStream_t st = ex.stream;
    const size_t base_mark = ex.arena->mark();
    u64 *recs = ex.alloc<u64>((u64) pl.n_buckets * Cfg::REC_CAP);
    kc_ull *cursor = reinterpret_cast<kc_ull *>(ex.alloc<u64>(pl.n_buckets));
    ex.fill_bytes(cursor, 0, (size_t) pl.n_buckets * 8);
    u32 *status = reinterpret_cast<u32 *>(cells + 3);
    kc_ull *m_cell = reinterpret_cast<kc_ull *>(cells + 2);
    constexpr u64 TILE = 256 * KC_EX_STRIP;
    {
        const u32 blocks = (u32) kc_div_up(n_bytes, TILE);
        const int n_parts = chunks && chunks->n > 1 && !chunks->waited ? chunks->n : 1;
        if (chunks && n_parts == 1) chunks->wait_all(st);
        for (int c = 0; c < n_parts; ++c) {
            u32 t0 = 0, t1 = blocks;
            if (n_parts > 1) {
                KC_CUDA(cudaStreamWaitEvent(st, chunks->ev[c], 0));
                t0 = (u32) std::min<u64>(blocks, (u64) c * (chunks->chunk_bytes / TILE));
                t1 = c == n_parts - 1 ? blocks : (u32) std::min<u64>(blocks, (u64) (c + 1) * (chunks->chunk_bytes / TILE));
            }
            if (t1 <= t0) continue;
            const u64 part_bytes = std::min<u64>(n_bytes, (u64) t1 * TILE) - (u64) t0 * TILE;
            CudaExec::Scope sc(ex, KP_KS_SCATTER0, part_bytes + part_bytes / 5);
            kc_sig_scan_launch(pl.m, t1 - t0, st, seq, n_bytes, k, pl.a, pl.n_buckets, Cfg::REC_CAP, Cfg::CAP, cursor, recs, flags,
                               (u32) (kc_div_up(n_bytes, (u64) 32) + 1), t0, m_cell, status, win_mask);
            ++ex.launches;
        }
        if (chunks) chunks->waited = true;
    }
    {
        const u32 fit = (u32) (dv.n_sm * (dv.occ > 0 ? dv.occ : 1));
        const u32 grid = pl.n_buckets < fit ? pl.n_buckets : fit;
        CudaExec::Scope sc(ex, KP_KS_RESOLVE, n_bytes / 5 + n_bytes * 10);
        kc_sig_resolve_kernel<L, false><<<grid, Cfg::THREADS, KK::smem_r(), st>>>(seq, n_bytes, k, complements ? 1 : 0, cursor, recs, pl.n_buckets,
                                                                                  kc_ksf_own_flags(flags), reinterpret_cast<kc_ull *>(cells), status);
        ++ex.launches;
        KC_CUDA(cudaGetLastError());
    }
    ex.arena->release(base_mark);
    return true;
}

#endif  // __CUDACC__

// Histogram-free k-mer set construction for the FLAGS-only regime (from-FASTA `compute` without -M): the same
// decode -> canonical k-mer -> MSD partition -> shared-memory hash resolve as kmerset.cuh (reference src/parser.h:22-141
// AddKMers / AddKMersWithFrequencies + khash), but planned from the input size alone so that NO counting pass and NO
// host synchronisation sits between the kernels.
//
// Why this is possible: FLAGS-only buckets by a bijective multiplicative hash of the k-mer word (kmer_scramble), so for
// any input made of mostly distinct k-mers the bucket sizes are Poisson-tight around M / ways.  Every bucket of every
// level therefore gets a FIXED slot of  mean + 8 sigma  items (the leaves: 1024, the capacity of the hash resolve) and
// a tile reserves its place inside a slot with one atomicAdd per digit.  What disappears against kmerset.cuh:
//   * kc_ks_hist0_kernel   — the sequence is decoded and the canonical k-mers are computed ONCE (they wait in shared
//                            memory while the tile's digit counts are scanned), not twice;
//   * kc_kv_hist_kernel    — no second read of the level-0 output (8L bytes per k-mer);
//   * two host read-backs (bucket lists / tile counts) and six scan / bookkeeping launches.
// Algorithmic HBM bytes per k-mer (L = 1): 1 + 12 (level 0) + 12 + 12 (level 1) + 12 (resolve) = 49 instead of 58.
//
// Inputs that break the assumption (30x read sets: every k-mer ~30 times, so sigma grows by sqrt(30); low-complexity
// sequence) overflow a slot.  Overflow is detected on the device (items beyond the slot are dropped, a status word is
// set) and reported with the same read-back that returns the counts; the caller then discards the flags and runs the
// exact histogram-based construction of kmerset.cuh.  Results never depend on which of the two ran.
#pragma once
#include "kmerset.cuh"

#include "ksf_plan.h"  // KSF_LEAF_CAP, KsfTuning, KsfPlan, kc_ksf_plan

// Host-to-device copy of the sequence in flight on another stream, cut into chunks of whole level-0 tiles: event c is recorded
// once bytes [0, (c + 1) * chunk_bytes) have landed.  The level-0 scatter of the fixed-slot construction waits per chunk, so
// the first partition pass runs behind the copy front instead of after the whole copy (kc_compute with host buffers).
struct InputChunks {
    int n = 0;
    u64 chunk_bytes = 0;     // a multiple of every KsCfg<L>::EX_TILE
    cudaEvent_t *ev = nullptr;
    bool waited = false;     // every chunk has been waited for on the compute stream
    void wait_all(cudaStream_t st) {
        if (n && !waited) KC_CUDA(cudaStreamWaitEvent(st, ev[n - 1], 0));
        waited = true;
    }
};

#ifdef __CUDACC__

// status[0] = 1: some slot overflowed (the flags are incomplete and must be discarded)
// ---- level 0: sequence -> fixed-slot buckets, one compute pass ----------------------------------------------------------
template <int L>
__global__ void __launch_bounds__(KsCfg<L>::EX_THREADS) kc_ksf_scatter0_kernel(const u8 *__restrict__ seq, u64 n_bytes, int k, int complements,
                                                                                int shift, int bits, u32 *bucket_cnt, u32 cap0,
                                                                                KWord<L> *__restrict__ keys, u32 *__restrict__ pos, u32 *status, u32 tile0,
                                                                                u32 *__restrict__ valid_flags = nullptr, u32 n_flag_words = 0) {
    constexpr int T = KsCfg<L>::EX_THREADS;
    constexpr int LOG_T = T == 256 ? 8 : (T == 128 ? 7 : 6);
    constexpr int R = 256 / T;
    constexpr int TILE = KsCfg<L>::EX_TILE;
    extern __shared__ __align__(16) unsigned char kc_smem_raw[];
    KWord<L> *stash = reinterpret_cast<KWord<L> *>(kc_smem_raw);  // slot j * T + t = window j of thread t (conflict-free)
    u16 *rk = reinterpret_cast<u16 *>(stash + TILE);               // rank of the slot's k-mer inside its digit, 0xFFFF = no k-mer
    u16 *perm = rk + TILE;                                         // digit-ordered position -> slot
    __shared__ u64 pk[KC_EX_HALO + T];
    __shared__ u32 vm[KC_EX_HALO + T];
    __shared__ u32 cnt[256];
    __shared__ u32 loff[256];
    __shared__ u64 dbase[256];  // slot index of the digit's first item of this tile, minus its staged position: at = dbase + q
    __shared__ u32 qlim[256];   // staged positions below this still fit into the digit's slot
    __shared__ u32 sw[T / 32];
    const i64 block_pos0 = (i64) (tile0 + blockIdx.x) * TILE;
    kc_tile_load<T>(seq, n_bytes, block_pos0, pk, vm);
    for (int i = threadIdx.x; i < 256; i += T) cnt[i] = 0;
#pragma unroll
    for (int j = 0; j < KC_EX_STRIP; ++j) rk[j * T + threadIdx.x] = 0xFFFFu;
    __syncthreads();
    const int widx = KC_EX_HALO + threadIdx.x;
    const u32 em = kc_strip_emit_mask(vm, widx, k);
    if (valid_flags) {  // clear-the-losers flags: every window starts as "first occurrence" (bit p & 31 of word p >> 5 = window END p)
        const u64 w = (u64) (block_pos0 >> 5) + threadIdx.x;
        if (em && w < n_flag_words) valid_flags[w] = __brev(em);
    }
    kc_strip_windows<L>(pk, widx, em, k, complements, [&](int j, const KWord<L> &c0) {
        const KWord<L> c = kmer_scramble(c0);
        const u32 slot = (u32) j * T + threadIdx.x;
        stash[slot] = c;
        rk[slot] = (u16) atomicAdd(&cnt[c.digit_top(shift, bits)], 1u);
    });
    __syncthreads();
    u32 total;
    {
        u32 v[R], c = 0;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            v[r] = cnt[threadIdx.x * R + r];
            c += v[r];
        }
        u32 p = kc_block_exclusive_scan<T>(c, &total, sw);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int i = threadIdx.x * R + r;
            loff[i] = p;
            u32 gb = 0;
            if (v[r]) gb = atomicAdd(&bucket_cnt[i], v[r]);  // reserve the tile's place inside the slot of bucket i
            const u32 room = gb < cap0 ? cap0 - gb : 0u;
            dbase[i] = (u64) i * cap0 + gb - p;
            qlim[i] = p + (v[r] < room ? v[r] : room);
            p += v[r];
        }
    }
    __syncthreads();
    if (total == 0) return;
    // perm[] shares the dynamic array with rk[] / stash[], so the compiler must keep every store behind the loads of the same
    // iteration and in front of the next one's (ncu r01h: 16 % of the kernel's samples on this one line).  Batches of 8 slots:
    // all ranks, all top limbs, all offsets, then the stores — the shared-memory round trips overlap.
#pragma unroll
    for (int j0 = 0; j0 < KC_EX_STRIP; j0 += 8) {
        u32 r[8], o[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) r[u] = rk[(u32) (j0 + u) * T + threadIdx.x];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            KWord<L> t = KWord<L>::zero();
            t.w[L - 1] = stash[(u32) (j0 + u) * T + threadIdx.x].w[L - 1];  // the digit lives in the top limb
            o[u] = t.digit_top(shift, bits);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) o[u] = loff[o[u]];
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (r[u] != 0xFFFFu) perm[o[u] + r[u]] = (u16) ((u32) (j0 + u) * T + threadIdx.x);
    }
    __syncthreads();
    bool over = false;
#pragma unroll 4
    for (u32 q = threadIdx.x; q < total; q += T) {
        const u32 slot = perm[q];
        const KWord<L> v = stash[slot];
        const u32 dg = v.digit_top(shift, bits);
        if (q < qlim[dg]) {
            const u64 at = dbase[dg] + q;
            keys[at] = v;
            pos[at] = (u32) block_pos0 + (slot & (T - 1)) * KC_EX_STRIP + (slot >> LOG_T);
        } else {
            over = true;
        }
    }
    if (over) status[0] = 1;
}

// L = 1 with TWO threads per 32-base strip: thread t handles the windows ending at bases [16 h, 16 h + 16) of strip
// t & 255, h = t >> 8.  Same tile, same shared memory, same output as kc_ksf_scatter0_kernel<1>, but 16 instead of 8 warps
// per CTA: the kernel is limited to 2 CTAs per SM by its 96 KB stash, and ncu showed it waiting on shared-memory
// round trips (short scoreboard) at 24 % occupancy.
__global__ void __launch_bounds__(512) kc_ksf_scatter0_split_kernel(const u8 *__restrict__ seq, u64 n_bytes, int k, int complements, int shift, int bits,
                                                                    u32 *bucket_cnt, u32 cap0, KWord<1> *__restrict__ keys, u32 *__restrict__ pos,
                                                                    u32 *status, u32 tile0) {
    constexpr int S = 256;   // strips per tile
    constexpr int T = 512;
    constexpr int TILE = KsCfg<1>::EX_TILE;
    extern __shared__ __align__(16) unsigned char kc_smem_raw[];
    u64 *stash = reinterpret_cast<u64 *>(kc_smem_raw);  // slot j * S + strip = window j of the strip (conflict-free)
    u16 *rk = reinterpret_cast<u16 *>(stash + TILE);
    u16 *perm = rk + TILE;
    __shared__ u64 pk[KC_EX_HALO + S];
    __shared__ u32 vm[KC_EX_HALO + S];
    __shared__ u32 cnt[256];
    __shared__ u32 loff[256];
    __shared__ u32 gbase[256];
    __shared__ u32 sw[T / 32];
    const i64 block_pos0 = (i64) (tile0 + blockIdx.x) * TILE;
    if (threadIdx.x < S) {
        kc_tile_load<S>(seq, n_bytes, block_pos0, pk, vm);
        cnt[threadIdx.x] = 0;
    }
    const u32 strip = threadIdx.x & (S - 1);
    const int j0 = (int) (threadIdx.x >> 8) * 16;
#pragma unroll
    for (int j = 0; j < 16; ++j) rk[(u32) (j0 + j) * S + strip] = 0xFFFFu;
    __syncthreads();
    const int widx = KC_EX_HALO + (int) strip;
    const u32 em = kc_strip_emit_mask(vm, widx, k);
    if ((em >> (16 - j0)) & 0xFFFFu) {  // bit 31 - j for base j: the half's 16 bits
        const u64 mine = pk[widx], prev = pk[widx - 1];
        const u64 mask = (1ULL << (2 * k)) - 1;
        const int top = 2 * (k - 1);
        // the k-1 bases before base j0 of the strip: the low end of (prev : mine) cut after base j0 - 1
        const u64 before = j0 ? ((prev << 32) | (mine >> 32)) : prev;
        KWord<1> pw;
        pw.w[0] = before & ((1ULL << top) - 1);
        u64 rcs = k > 1 ? kmer_reverse_complement(pw, k - 1).w[0] : 0;
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) {
            const int j = j0 + jj;
            const int sh = 2 * (31 - j);
            u64 fwd = mine >> sh;
            if (j < 31) fwd |= prev << (2 * (j + 1));
            fwd &= mask;
            const u64 c = (mine >> sh) & 3;
            const u64 rcf = rcs | ((3 ^ c) << top);
            rcs = rcf >> 2;
            if ((em >> (31 - j)) & 1) {
                KWord<1> canon;
                canon.w[0] = (!complements || fwd < rcf) ? fwd : rcf;
                const KWord<1> cs = kmer_scramble(canon);
                const u32 slot = (u32) j * S + strip;
                stash[slot] = cs.w[0];
                rk[slot] = (u16) atomicAdd(&cnt[cs.digit_top(shift, bits)], 1u);
            }
        }
    }
    __syncthreads();
    u32 total;
    {
        const u32 c = threadIdx.x < 256 ? cnt[threadIdx.x] : 0u;
        const u32 p = kc_block_exclusive_scan<T>(c, &total, sw);
        if (threadIdx.x < 256) {
            loff[threadIdx.x] = p;
            if (c) gbase[threadIdx.x] = atomicAdd(&bucket_cnt[threadIdx.x], c);
        }
    }
    __syncthreads();
    if (total == 0) return;
    for (u32 slot = threadIdx.x; slot < (u32) TILE; slot += T) {
        const u32 r = rk[slot];
        if (r != 0xFFFFu) {
            KWord<1> v;
            v.w[0] = stash[slot];
            perm[loff[v.digit_top(shift, bits)] + r] = (u16) slot;
        }
    }
    __syncthreads();
    bool over = false;
    for (u32 q = threadIdx.x; q < total; q += T) {
        const u32 slot = perm[q];
        KWord<1> v;
        v.w[0] = stash[slot];
        const u32 dg = v.digit_top(shift, bits);
        const u32 idx = gbase[dg] + (q - loff[dg]);
        if (idx < cap0) {
            const u64 at = (u64) dg * cap0 + idx;
            keys[at] = v;
            pos[at] = (u32) ((u64) block_pos0 + (slot & (S - 1)) * KC_EX_STRIP + (slot >> 8));
        } else {
            over = true;
        }
    }
    if (over) status[0] = 1;
}

// Between two levels: clamp the fill counts of the nP parent slots, count their tiles (tile_count[nP] = 0 for the scan
// that follows) and, after level 0, add up M.
__global__ void __launch_bounds__(256) kc_ksf_prep_kernel(const u32 *cnt, u32 nP, u32 capP, u32 tile, u32 *size_out, u32 *tile_count, u32 *status,
                                                          kc_ull *m_cell) {
    const u32 p = blockIdx.x * 256 + threadIdx.x;
    if (p > nP) return;
    if (p == nP) {
        tile_count[p] = 0;
        return;
    }
    u32 c = cnt[p];
    if (m_cell && c) atomicAdd(m_cell, (kc_ull) (c < capP ? c : capP));
    if (c > capP) {
        status[0] = 1;
        c = capP;
    }
    size_out[p] = c;
    tile_count[p] = (c + tile - 1) / tile;
}

// ---- levels >= 1: fixed-slot parents -> fixed-slot children (kc_kv_scatter_kernel without the counting pass) ---------------
template <int L, int TILE, int MINB>
__global__ void __launch_bounds__(256, MINB) kc_ksf_scatter_kernel(const KWord<L> *__restrict__ ksrc, const u32 *__restrict__ psrc, KWord<L> *__restrict__ kdst,
                                                             u32 *__restrict__ pdst, const u32 *__restrict__ P_size, const u32 *__restrict__ tile_prefix,
                                                             u32 nP, u64 capP, u32 tiles_per_cta, int shift, int bits, u32 *C_cnt, u32 capC, u32 *status,
                                                             const u64 *__restrict__ P_off = nullptr) {
    constexpr int ITEMS = TILE / 256;
    extern __shared__ __align__(16) unsigned char kc_smem_raw[];
    KWord<L> *stage_k = reinterpret_cast<KWord<L> *>(kc_smem_raw);
    u32 *stage_p = reinterpret_cast<u32 *>(stage_k + TILE);
    u16 *rk = reinterpret_cast<u16 *>(stage_p + TILE);
    __shared__ u32 cnt[256];
    __shared__ u32 loff[256];
    __shared__ u32 gbase[256];
    __shared__ u32 sw[8];
    const u32 n_tiles = tile_prefix[nP];
    const u32 t0 = blockIdx.x * tiles_per_cta;
    const u32 t1 = min(n_tiles, t0 + tiles_per_cta);
    if (t0 >= t1) return;
    u32 b = kc_upper_bound_u32(tile_prefix, nP + 1, t0) - 1;
    cnt[threadIdx.x] = 0;
    bool over = false;
    __syncthreads();
    for (u32 t = t0; t < t1; ++t) {
        while (t >= tile_prefix[b + 1]) ++b;
        const u64 pbase = P_off ? P_off[b] : (u64) b * capP;  // P_off: parents laid out back to back (multi-GPU receive buffer)
        const KWord<L> *src = ksrc + pbase;
        const u32 *ps = psrc + pbase;
        const u32 start = (t - tile_prefix[b]) * TILE;
        const u32 n_here = min((u32) TILE, P_size[b] - start);
        KWord<L> item[ITEMS];
        u32 pay[ITEMS];
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            const u32 i = threadIdx.x + j * 256;
            if (i < n_here) item[j] = src[start + i];
        }
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            const u32 i = threadIdx.x + j * 256;
            if (i < n_here) pay[j] = ps[start + i];
        }
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            const u32 i = threadIdx.x + j * 256;
            if (i < n_here) rk[i] = (u16) atomicAdd(&cnt[item[j].digit_top(shift, bits)], 1u);
        }
        __syncthreads();
        const u32 c = cnt[threadIdx.x];
        u32 total;
        const u32 p = kc_block_exclusive_scan<256>(c, &total, sw);
        loff[threadIdx.x] = p;
        if (c) gbase[threadIdx.x] = atomicAdd(&C_cnt[((u64) b << bits) + threadIdx.x], c);
        cnt[threadIdx.x] = 0;
        __syncthreads();
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            const u32 i = threadIdx.x + j * 256;
            if (i < n_here) {
                const u32 q = loff[item[j].digit_top(shift, bits)] + rk[i];
                stage_k[q] = item[j];
                stage_p[q] = pay[j];
            }
        }
        __syncthreads();
        for (u32 q = threadIdx.x; q < n_here; q += 256) {
            const KWord<L> v = stage_k[q];
            const u32 dg = v.digit_top(shift, bits);
            const u32 idx = gbase[dg] + (q - loff[dg]);
            if (idx < capC) {
                const u64 at = (((u64) b << bits) + dg) * capC + idx;
                kdst[at] = v;
                pdst[at] = stage_p[q];
            } else {
                over = true;
            }
        }
        __syncthreads();
    }
    if (over) status[0] = 1;
}

// Leaf slots -> the bucket list kc_ks_resolve_hash_kernel walks.
__global__ void __launch_bounds__(256) kc_ksf_leaf_kernel(const u32 *cnt, u64 n_leaf, u32 cap, u8 parity, SortBucket *small, u32 *status) {
    const u64 c = (u64) blockIdx.x * 256 + threadIdx.x;
    if (c >= n_leaf) return;
    u32 s = cnt[c];
    if (s > cap) {
        status[0] = 1;
        s = cap;
    }
    SortBucket d;
    d.off = c * cap;
    d.size = s;
    d.rem = 0;
    d.parity = parity;
    d.bits = 0;
    small[c] = d;
}

// ---- leaf resolve: exact dedup of one fixed slot through shared-memory tables, two barriers per bucket ------------------------
// Persistent CTAs walk the leaf slots with stride gridDim.x.  The slot of bucket n + 1 streams into the second staging
// buffer with cp.async while bucket n is resolved, so no thread ever waits on HBM between the barriers.
//   A  every item writes its index into T1[h1(key)] with a plain store (some writer wins);
//   -- barrier --
//   B  the winner of a slot represents its key.  An item that finds an EQUAL key there is a duplicate: it folds its
//      position into the winner's with atomicMin (and bumps the occurrence count for -z).  An item that finds a
//      DIFFERENT key (~16 % at load 0.37) inserts itself into T2 with atomicCAS + linear probing, folding into an
//      equal key if it meets one.  All copies of a key take the same route, so they always meet;
//   -- barrier --
//   C  every representative with >= min_count occurrences sets the bit of its smallest position.
// Against kc_ks_resolve_hash_kernel (up to six write-then-verify rounds, each ending in a barrier + vote): 2 barriers
// instead of ~5 per bucket, no register staging (64 -> 40 registers), twice the resident CTAs.
KC_D void kc_cp_async16(void *smem_dst, const void *gmem_src) {
    const u32 d = (u32) __cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
KC_D void kc_cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
KC_D void kc_cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

KC_D void kc_flag_clear(u32 *flags, u32 pos) { atomicAnd(&flags[pos >> 5], ~(1u << (pos & 31))); }

// CLEAR = false: `flags` arrives zeroed and every kept k-mer sets the bit of its smallest position (M scattered REDs).
// CLEAR = true : `flags` arrives with the bit of every valid window set (written by level 0, one coalesced word per strip);
//                every atomicMin that folds a duplicate knocks out exactly one position — the larger of the two it compared —
//                so only the DUPLICATES cost a scattered RED (a genome: ~0 of them) and the kept bits are never touched.
template <int L, bool COUNTED, bool CLEAR = false>
__global__ void __launch_bounds__(256) kc_ksf_resolve_kernel(const KWord<L> *__restrict__ keys, const u32 *__restrict__ pos, const u32 *__restrict__ cnt,
                                                             u32 n_leaf, u32 *flags, u32 min_count, kc_ull *n_unique, u32 *status) {
    constexpr u32 CAP = KSF_LEAF_CAP;
    constexpr u32 T1N = 2 * CAP, T2N = CAP;
    constexpr int KPC = 16 / (int) sizeof(KWord<L>) > 0 ? 16 / (int) sizeof(KWord<L>) : 1;  // keys per 16-byte chunk (L = 1: 2)
    constexpr int CPK = (int) sizeof(KWord<L>) / 16 > 0 ? (int) sizeof(KWord<L>) / 16 : 1;  // 16-byte chunks per key (L = 4: 2)
    extern __shared__ __align__(16) unsigned char kc_smem_raw[];
    KWord<L> *sk0 = reinterpret_cast<KWord<L> *>(kc_smem_raw);
    u32 *sp0 = reinterpret_cast<u32 *>(sk0 + 2 * CAP);
    u32 *T1 = sp0 + 2 * CAP;
    u32 *T2 = T1 + T1N;
    u32 *occ = T2 + T2N;  // COUNTED only
    const u32 stride = gridDim.x;
    u32 c = blockIdx.x;
    if (c >= n_leaf) return;
    auto fetch = [&](u32 bucket, u32 size, int buf) {
        const char *gk = reinterpret_cast<const char *>(keys + (u64) bucket * CAP);
        const char *gp = reinterpret_cast<const char *>(pos + (u64) bucket * CAP);
        char *dk = reinterpret_cast<char *>(sk0 + (u32) buf * CAP);
        char *dp = reinterpret_cast<char *>(sp0 + (u32) buf * CAP);
        // a thread copies whole units (L = 1: a pair of keys, otherwise one key), so the keys it hashes in phase A are
        // the ones it fetched itself
        const u32 n_units = (size + KPC - 1) / KPC, pchunks = (size * 4 + 15) / 16;
        for (u32 u = threadIdx.x; u < n_units; u += 256) {
#pragma unroll
            for (int cc = 0; cc < CPK; ++cc) kc_cp_async16(dk + 16 * (u * CPK + cc), gk + 16 * (u * CPK + cc));
        }
        for (u32 q = threadIdx.x; q < pchunks; q += 256) kc_cp_async16(dp + 16 * q, gp + 16 * q);
        kc_cp_async_commit();
    };
    u32 size = cnt[c];
    if (size > CAP) {
        size = CAP;
        if (threadIdx.x == 0) status[0] = 1;
    }
    fetch(c, size, 0);
    int buf = 0;
    u32 kept = 0;
    while (true) {
        const u32 cn = c + stride;
        u32 size_n = 0;
        if (cn < n_leaf) {
            size_n = cnt[cn];
            if (size_n > CAP) {
                size_n = CAP;
                if (threadIdx.x == 0) status[0] = 1;
            }
        }
        KWord<L> *sk = sk0 + (u32) buf * CAP;
        u32 *sp = sp0 + (u32) buf * CAP;
        kc_cp_async_wait_all();  // the thread's own chunks of bucket c have landed
        // A: own items = the keys of the chunks this thread copied itself (visible without a barrier)
        const u32 n_chunk_items = (size + KPC - 1) / KPC;  // L = 1: pairs of keys; L >= 2: single keys
        {
            const uint4 e = make_uint4(KC_NONE, KC_NONE, KC_NONE, KC_NONE);
            reinterpret_cast<uint4 *>(T2)[threadIdx.x] = e;  // T2N = 1024 slots = 256 x 16 bytes
        }
        for (u32 q = threadIdx.x; q < n_chunk_items; q += 256) {
#pragma unroll
            for (int e = 0; e < KPC; ++e) {
                const u32 i = q * KPC + e;
                if (i < size) {
                    u64 h = 0;
#pragma unroll
                    for (int w = 0; w < L; ++w) h = (h ^ sk[i].w[w]) * 0xD6E8FEB86659FD93ULL;
                    T1[h >> 53] = i;
                    if (COUNTED) occ[i] = 1;
                }
            }
        }
        __syncthreads();
        if (cn < n_leaf) fetch(cn, size_n, buf ^ 1);  // every thread is past phase C of the bucket that used that buffer
        // B
        u32 rep = 0;  // bit (2 m + e): item e of the thread's m-th chunk represents its key
        {
            u32 m = 0;
            for (u32 q = threadIdx.x; q < n_chunk_items; q += 256, ++m) {
#pragma unroll
                for (int e = 0; e < KPC; ++e) {
                    const u32 i = q * KPC + e;
                    if (i >= size) continue;
                    const KWord<L> key = sk[i];
                    u64 h = 0;
#pragma unroll
                    for (int w = 0; w < L; ++w) h = (h ^ key.w[w]) * 0xD6E8FEB86659FD93ULL;
                    const u32 o = T1[h >> 53];
                    if (o == i) {
                        rep |= 1u << (m * KPC + e);
                    } else if (sk[o] == key) {
                        const u32 mine = sp[i];
                        const u32 was = atomicMin(&sp[o], mine);
                        if (CLEAR) kc_flag_clear(flags, was > mine ? was : mine);
                        if (COUNTED) atomicAdd(&occ[o], 1u);
                    } else {
                        u32 s = (u32) (h >> 43) & (T2N - 1);
                        while (true) {
                            const u32 old = atomicCAS(&T2[s], KC_NONE, i);
                            if (old == KC_NONE) {
                                rep |= 1u << (m * KPC + e);
                                break;
                            }
                            if (sk[old] == key) {
                                const u32 mine = sp[i];
                                const u32 was = atomicMin(&sp[old], mine);
                                if (CLEAR) kc_flag_clear(flags, was > mine ? was : mine);
                                if (COUNTED) atomicAdd(&occ[old], 1u);
                                break;
                            }
                            s = (s + 1) & (T2N - 1);
                        }
                    }
                }
            }
        }
        __syncthreads();
        // C
        {
            u32 m = 0;
            for (u32 q = threadIdx.x; q < n_chunk_items; q += 256, ++m) {
#pragma unroll
                for (int e = 0; e < KPC; ++e) {
                    const u32 i = q * KPC + e;
                    if (!((rep >> (m * KPC + e)) & 1u)) continue;
                    if (!COUNTED || occ[i] >= min_count) {
                        if (!CLEAR) kc_flag_set(flags, sp[i]);
                        ++kept;
                    } else if (CLEAR) {
                        kc_flag_clear(flags, sp[i]);  // too few occurrences: the surviving (smallest) position goes as well
                    }
                }
            }
        }
        if (cn >= n_leaf) break;
        c = cn;
        size = size_n;
        buf ^= 1;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) kept += __shfl_down_sync(0xFFFFFFFFu, kept, o);
    if ((threadIdx.x & 31) == 0 && kept) atomicAdd(n_unique, (kc_ull) kept);
}

// ---- leaf resolve, one barrier per leaf (-z 1, clear-the-losers flags) ---------------------------------------------------------
// With clear-the-losers flags phase C of kc_ksf_resolve_kernel has nothing left to write, so the barrier in front of it only
// protects the tables from the next leaf's phase A.  Two sets of tables (T1 as u16 indices, T2 as u32 for the CAS) alternate
// instead: leaf n + 1 fills the set that leaf n - 1 used, and every thread is past leaf n - 1 once it has crossed the barrier
// of leaf n.
// STAGES staging buffers form a ring: while leaf n is resolved, leaves n + 1 .. n + STAGES - 1 are in flight (cp.async groups,
// one per leaf).  With two stages the kernel ran at 2.5 TB/s = exactly the bytes in flight (5 CTAs x 9 KB per SM) over the
// ~2.7 us a leaf took, i.e. it was bound by DRAM latency through Little's law, not by its instructions.
// Shared memory per CTA (L = 1): STAGES x 12 KB + 16 KB of tables.
template <int N> KC_D void kc_cp_async_wait_group() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

template <int L, int STAGES>
__global__ void __launch_bounds__(256) kc_ksf_resolve1_kernel(const KWord<L> *__restrict__ keys, const u32 *__restrict__ pos, const u32 *__restrict__ cnt,
                                                              u32 n_leaf, u32 *flags, kc_ull *n_unique, u32 *status) {
    constexpr u32 CAP = KSF_LEAF_CAP;
    constexpr u32 T1N = 2 * CAP, T2N = CAP;
    constexpr int KPC = 16 / (int) sizeof(KWord<L>) > 0 ? 16 / (int) sizeof(KWord<L>) : 1;
    constexpr int CPK = (int) sizeof(KWord<L>) / 16 > 0 ? (int) sizeof(KWord<L>) / 16 : 1;
    extern __shared__ __align__(16) unsigned char kc_smem_raw[];
    KWord<L> *sk0 = reinterpret_cast<KWord<L> *>(kc_smem_raw);
    u32 *sp0 = reinterpret_cast<u32 *>(sk0 + STAGES * CAP);
    u32 *T2a = sp0 + STAGES * CAP;                             // [2][T2N]
    u16 *T1a = reinterpret_cast<u16 *>(T2a + 2 * T2N);         // [2][T1N]
    const u64 stride = gridDim.x;
    u64 c = blockIdx.x;
    if (c >= n_leaf) return;
    auto leaf_size = [&](u64 leaf) -> u32 {
        if (leaf >= n_leaf) return 0u;
        u32 sz = cnt[leaf];
        if (sz > CAP) {
            sz = CAP;
            if (threadIdx.x == 0) status[0] = 1;
        }
        return sz;
    };
    // one cp.async group per leaf, empty when the CTA has run out of leaves (keeps the group count uniform)
    auto fetch = [&](u64 bucket, u32 size, int buf) {
        if (bucket < n_leaf) {
            const char *gk = reinterpret_cast<const char *>(keys + bucket * CAP);
            const char *gp = reinterpret_cast<const char *>(pos + bucket * CAP);
            char *dk = reinterpret_cast<char *>(sk0 + (u32) buf * CAP);
            char *dp = reinterpret_cast<char *>(sp0 + (u32) buf * CAP);
            // a thread copies whole units (L = 1: a pair of keys, otherwise one key), so the keys it hashes in phase A are the
            // ones it fetched itself
            const u32 n_units = (size + KPC - 1) / KPC, pchunks = (size * 4 + 15) / 16;
            for (u32 u = threadIdx.x; u < n_units; u += 256) {
#pragma unroll
                for (int cc = 0; cc < CPK; ++cc) kc_cp_async16(dk + 16 * (u * CPK + cc), gk + 16 * (u * CPK + cc));
            }
            for (u32 q = threadIdx.x; q < pchunks; q += 256) kc_cp_async16(dp + 16 * q, gp + 16 * q);
        }
        kc_cp_async_commit();
    };
    u32 sz[STAGES];  // sz[i] = size of leaf c + i * stride
#pragma unroll
    for (int i = 0; i < STAGES - 1; ++i) {
        sz[i] = leaf_size(c + (u64) i * stride);
        fetch(c + (u64) i * stride, sz[i], i);
    }
    int buf = 0, par = 0;
    u32 kept = 0;
    while (true) {
        const u64 c_far = c + (u64) (STAGES - 1) * stride;
        sz[STAGES - 1] = leaf_size(c_far);
        const u32 size = sz[0];
        KWord<L> *sk = sk0 + (u32) buf * CAP;
        u32 *sp = sp0 + (u32) buf * CAP;
        u32 *T2 = T2a + (u32) par * T2N;
        u16 *T1 = T1a + (u32) par * T1N;
        kc_cp_async_wait_group<STAGES - 2>();  // the thread's own chunks of leaf c have landed
        const u32 n_chunk_items = (size + KPC - 1) / KPC;
        reinterpret_cast<uint4 *>(T2)[threadIdx.x] = make_uint4(KC_NONE, KC_NONE, KC_NONE, KC_NONE);  // T2N = 256 x 4 slots
        // A: some item of every key group wins the group's T1 slot (plain 16-bit stores)
        for (u32 q = threadIdx.x; q < n_chunk_items; q += 256) {
#pragma unroll
            for (int e = 0; e < KPC; ++e) {
                const u32 i = q * KPC + e;
                if (i < size) {
                    u64 h = 0;
#pragma unroll
                    for (int w = 0; w < L; ++w) h = (h ^ sk[i].w[w]) * 0xD6E8FEB86659FD93ULL;
                    T1[h >> 53] = (u16) i;
                }
            }
        }
        __syncthreads();
        {   // every thread is past phase B of the leaf that used the ring slot behind the newest one
            int far = buf + STAGES - 1;
            if (far >= STAGES) far -= STAGES;
            fetch(c_far, sz[STAGES - 1], far);
        }
        // B: winners represent their key; a duplicate folds its position and clears the larger of the two
        for (u32 q = threadIdx.x; q < n_chunk_items; q += 256) {
#pragma unroll
            for (int e = 0; e < KPC; ++e) {
                const u32 i = q * KPC + e;
                if (i >= size) continue;
                const KWord<L> key = sk[i];
                u64 h = 0;
#pragma unroll
                for (int w = 0; w < L; ++w) h = (h ^ key.w[w]) * 0xD6E8FEB86659FD93ULL;
                const u32 o = T1[h >> 53];
                if (o == i) {
                    ++kept;
                } else if (sk[o] == key) {
                    const u32 mine = sp[i];
                    const u32 was = atomicMin(&sp[o], mine);
                    kc_flag_clear(flags, was > mine ? was : mine);
                } else {
                    u32 s = (u32) (h >> 43) & (T2N - 1);
                    while (true) {
                        const u32 old = atomicCAS(&T2[s], KC_NONE, i);
                        if (old == KC_NONE) {
                            ++kept;
                            break;
                        }
                        if (sk[old] == key) {
                            const u32 mine = sp[i];
                            const u32 was = atomicMin(&sp[old], mine);
                            kc_flag_clear(flags, was > mine ? was : mine);
                            break;
                        }
                        s = (s + 1) & (T2N - 1);
                    }
                }
            }
        }
        if (c + stride >= n_leaf) break;
        c += stride;
#pragma unroll
        for (int i = 0; i < STAGES - 1; ++i) sz[i] = sz[i + 1];
        buf = buf + 1 == STAGES ? 0 : buf + 1;
        par ^= 1;
    }
    kc_cp_async_wait_group<0>();  // only empty groups are left
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) kept += __shfl_down_sync(0xFFFFFFFFu, kept, o);
    if ((threadIdx.x & 31) == 0 && kept) atomicAdd(n_unique, (kc_ull) kept);
}

// ---- one-barrier leaf resolve with the thread's items staged in registers ------------------------------------------------------
// The stage sweep of kc_ksf_resolve1_kernel (profiles/r01h_variant_sweep2.json) showed its time going with 1 / (CTAs per SM):
// 0.235 ms at 5 CTAs (2 stages), 0.270 at 4 (3 stages), 0.324 at 3 (4 stages) — the kernel is bound by the latency of one
// leaf's dependent chain (LDS key -> hash -> STS/LDS table -> compare, item after item), not by DRAM.  Here a thread loads
// ALL its keys first, hashes them together, then issues the table accesses together, so the chain is paid once per phase
// instead of once per item; key and hash stay in registers between the phases.  THREADS = 512 halves the items per thread.
// BAL (measured next, ncu r01h: 30 % of the samples wait at the barrier, 6 % on the load of the next leaf's size): with 16-byte
// copy units a leaf of ~763 u64 keys gives threads 0..125 four items and the others two, so half the warps idle at the barrier;
// 8-byte units (cp.async.ca) give every thread three.  The size of leaf n + 2 is loaded one iteration before it is needed.
KC_D void kc_cp_async8(void *smem_dst, const void *gmem_src) {
    const u32 d = (u32) __cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gmem_src) : "memory");
}

template <int L, int THREADS, bool BAL = false>
__global__ void __launch_bounds__(THREADS) kc_ksf_resolve2_kernel(const KWord<L> *__restrict__ keys, const u32 *__restrict__ pos, const u32 *__restrict__ cnt,
                                                                  u32 n_leaf, u32 *flags, kc_ull *n_unique, u32 *status) {
    constexpr u32 CAP = KSF_LEAF_CAP;
    constexpr u32 T1N = 2 * CAP, T2N = CAP;
    constexpr int KPC = (BAL && L == 1) ? 1 : (16 / (int) sizeof(KWord<L>) > 0 ? 16 / (int) sizeof(KWord<L>) : 1);
    constexpr int CPK = (int) sizeof(KWord<L>) / 16 > 0 ? (int) sizeof(KWord<L>) / 16 : 1;
    constexpr int UNITS = (int) CAP / KPC / THREADS;  // copy units (and hash rounds) per thread
    constexpr int NI = UNITS * KPC;                   // items per thread
    static_assert(UNITS >= 1 && UNITS * KPC * THREADS == (int) CAP, "CAP must be a whole number of units per thread");
    extern __shared__ __align__(16) unsigned char kc_smem_raw[];
    KWord<L> *sk0 = reinterpret_cast<KWord<L> *>(kc_smem_raw);
    u32 *sp0 = reinterpret_cast<u32 *>(sk0 + 2 * CAP);
    u32 *T2a = sp0 + 2 * CAP;                                  // [2][T2N]
    u16 *T1a = reinterpret_cast<u16 *>(T2a + 2 * T2N);         // [2][T1N]
    const u64 stride = gridDim.x;
    u64 c = blockIdx.x;
    if (c >= n_leaf) return;
    auto leaf_size = [&](u64 leaf) -> u32 {
        if (leaf >= n_leaf) return 0u;
        u32 sz = cnt[leaf];
        if (sz > CAP) {
            sz = CAP;
            if (threadIdx.x == 0) status[0] = 1;
        }
        return sz;
    };
    auto fetch = [&](u64 bucket, u32 size, int buf) {
        if (bucket < n_leaf) {
            const char *gk = reinterpret_cast<const char *>(keys + bucket * CAP);
            const char *gp = reinterpret_cast<const char *>(pos + bucket * CAP);
            char *dk = reinterpret_cast<char *>(sk0 + (u32) buf * CAP);
            char *dp = reinterpret_cast<char *>(sp0 + (u32) buf * CAP);
            const u32 n_units = (size + KPC - 1) / KPC, pchunks = (size * 4 + 15) / 16;
#pragma unroll
            for (int m = 0; m < UNITS; ++m) {  // unit u holds items u * KPC .. u * KPC + KPC - 1: the ones this thread hashes
                const u32 u = threadIdx.x + (u32) m * THREADS;
                if (u < n_units) {
                    if constexpr (BAL && L == 1) {
                        kc_cp_async8(dk + 8 * u, gk + 8 * u);
                    } else {
#pragma unroll
                        for (int cc = 0; cc < CPK; ++cc) kc_cp_async16(dk + 16 * (u * CPK + cc), gk + 16 * (u * CPK + cc));
                    }
                }
            }
            for (u32 q = threadIdx.x; q < pchunks; q += THREADS) kc_cp_async16(dp + 16 * q, gp + 16 * q);
        }
        kc_cp_async_commit();
    };
    u32 size = leaf_size(c);
    u32 size_n = leaf_size(c + stride);
    fetch(c, size, 0);
    int buf = 0;
    u32 kept = 0;
    while (true) {
        const u64 cn = c + stride;
        u32 raw_nn = 0;  // size of leaf c + 2 * stride: loaded here, first used at the bottom of the loop
        if (BAL) {
            if (cn + stride < n_leaf) raw_nn = cnt[cn + stride];
        }
        KWord<L> *sk = sk0 + (u32) buf * CAP;
        u32 *sp = sp0 + (u32) buf * CAP;
        u32 *T2 = T2a + (u32) buf * T2N;
        u16 *T1 = T1a + (u32) buf * T1N;
        kc_cp_async_wait_group<0>();  // the thread's own key units of leaf c have landed
        if (threadIdx.x < 256) reinterpret_cast<uint4 *>(T2)[threadIdx.x] = make_uint4(KC_NONE, KC_NONE, KC_NONE, KC_NONE);  // T2N = 256 x 4 slots
        // A: all keys, then all hashes, then all table stores
        KWord<L> key[NI];
        u32 h1[NI], h2[NI];
#pragma unroll
        for (int m = 0; m < NI; ++m) {
            const u32 i = (threadIdx.x + (u32) (m / KPC) * THREADS) * KPC + (u32) (m % KPC);
            if (i < size) key[m] = sk[i];
        }
#pragma unroll
        for (int m = 0; m < NI; ++m) {
            u64 h = 0;
#pragma unroll
            for (int w = 0; w < L; ++w) h = (h ^ key[m].w[w]) * 0xD6E8FEB86659FD93ULL;
            h1[m] = (u32) (h >> 53);
            h2[m] = (u32) (h >> 43) & (T2N - 1);
        }
#pragma unroll
        for (int m = 0; m < NI; ++m) {
            const u32 i = (threadIdx.x + (u32) (m / KPC) * THREADS) * KPC + (u32) (m % KPC);
            if (i < size) T1[h1[m]] = (u16) i;
        }
        __syncthreads();
        fetch(cn, size_n, buf ^ 1);  // every thread is past phase B of the leaf that used that buffer
        // B: all table reads first; winners are done, the others compare against their slot's winner
        u32 o[NI];
#pragma unroll
        for (int m = 0; m < NI; ++m) {
            const u32 i = (threadIdx.x + (u32) (m / KPC) * THREADS) * KPC + (u32) (m % KPC);
            o[m] = i < size ? (u32) T1[h1[m]] : i;
        }
#pragma unroll
        for (int m = 0; m < NI; ++m) {
            const u32 i = (threadIdx.x + (u32) (m / KPC) * THREADS) * KPC + (u32) (m % KPC);
            if (i >= size) continue;
            if (o[m] == i) {
                ++kept;
            } else if (sk[o[m]] == key[m]) {
                const u32 mine = sp[i];
                const u32 was = atomicMin(&sp[o[m]], mine);
                kc_flag_clear(flags, was > mine ? was : mine);
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
                        kc_flag_clear(flags, was > mine ? was : mine);
                        break;
                    }
                    s = (s + 1) & (T2N - 1);
                }
            }
        }
        if (cn >= n_leaf) break;
        c = cn;
        size = size_n;
        if (BAL) {
            if (raw_nn > CAP) {
                raw_nn = CAP;
                if (threadIdx.x == 0) status[0] = 1;
            }
            size_n = raw_nn;
        } else {
            size_n = leaf_size(c + stride);
        }
        buf ^= 1;
    }
    kc_cp_async_wait_group<0>();
#pragma unroll
    for (int o2 = 16; o2 > 0; o2 >>= 1) kept += __shfl_down_sync(0xFFFFFFFFu, kept, o2);
    if ((threadIdx.x & 31) == 0 && kept) atomicAdd(n_unique, (kc_ull) kept);
}

// ---- levels >= 1 with the next tile in flight ---------------------------------------------------------------------------------
// kc_ksf_scatter_kernel issues a tile's loads and then waits for them (ncu: long scoreboard, 47 % of the HBM peak).  Here the
// tile after the current one streams into a shared-memory input buffer with cp.async while the current tile is ranked,
// staged and written out: a thread takes its items out of the input buffer into registers, and the barrier that ends the
// ranking phase frees the buffer for the next copy.  Tiles are contiguous in their parent slot, slots are 128-byte aligned.
template <int L, int TILE, int MINB, int THREADS = 256>
__global__ void __launch_bounds__(THREADS, MINB) kc_ksf_scatter_pf_kernel(const KWord<L> *__restrict__ ksrc, const u32 *__restrict__ psrc, KWord<L> *__restrict__ kdst,
                                                                u32 *__restrict__ pdst, const u32 *__restrict__ P_size, const u32 *__restrict__ tile_prefix,
                                                                u32 nP, u64 capP, u32 tiles_per_cta, int shift, int bits, u32 *C_cnt, u32 capC, u32 *status) {
    constexpr int ITEMS = TILE / THREADS;
    extern __shared__ __align__(16) unsigned char kc_smem_raw[];
    KWord<L> *in_k = reinterpret_cast<KWord<L> *>(kc_smem_raw);
    KWord<L> *stage_k = in_k + TILE;
    u32 *in_p = reinterpret_cast<u32 *>(stage_k + TILE);
    u32 *stage_p = in_p + TILE;
    u16 *rk = reinterpret_cast<u16 *>(stage_p + TILE);
    __shared__ u32 cnt[256];
    __shared__ u32 loff[256];
    __shared__ u64 dbase[256];  // slot index of the digit's first item of this tile, minus its staged position: at = dbase + q
    __shared__ u32 qlim[256];   // staged positions below this still fit into the child's slot
    __shared__ u32 sw[THREADS / 32];
    const u32 n_tiles = tile_prefix[nP];
    const u32 t0 = blockIdx.x * tiles_per_cta;
    const u32 t1 = min(n_tiles, t0 + tiles_per_cta);
    if (t0 >= t1) return;
    u32 b = kc_upper_bound_u32(tile_prefix, nP + 1, t0) - 1;
    if (threadIdx.x < 256) cnt[threadIdx.x] = 0;
    bool over = false;
    // tile t of the CTA -> (parent bucket, first item inside the source arrays, items)
    auto describe = [&](u32 t, u32 &bb, u64 &first, u32 &n_here) {
        while (t >= tile_prefix[bb + 1]) ++bb;
        const u32 start = (t - tile_prefix[bb]) * TILE;
        n_here = min((u32) TILE, P_size[bb] - start);
        first = (u64) bb * capP + start;
    };
    auto prefetch = [&](u64 first, u32 n_here) {
        const char *gk = reinterpret_cast<const char *>(ksrc + first);
        const char *gp = reinterpret_cast<const char *>(psrc + first);
        const u32 kch = (n_here * (u32) sizeof(KWord<L>) + 15) / 16, pch = (n_here * 4 + 15) / 16;
        for (u32 q = threadIdx.x; q < kch; q += THREADS) kc_cp_async16(reinterpret_cast<char *>(in_k) + 16 * q, gk + 16 * q);
        for (u32 q = threadIdx.x; q < pch; q += THREADS) kc_cp_async16(reinterpret_cast<char *>(in_p) + 16 * q, gp + 16 * q);
        kc_cp_async_commit();
    };
    u64 first;
    u32 n_here;
    describe(t0, b, first, n_here);
    prefetch(first, n_here);
    for (u32 t = t0; t < t1; ++t) {
        kc_cp_async_wait_all();
        __syncthreads();  // tile t is complete in the input buffer; the write-out of tile t - 1 has left the staging buffers
        KWord<L> item[ITEMS];
        u32 pay[ITEMS];
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            const u32 i = threadIdx.x + j * THREADS;
            if (i < n_here) {
                item[j] = in_k[i];
                pay[j] = in_p[i];
            }
        }
        // all atomics first, then the rank stores (a store behind every atomic would wait for its result before the next
        // atomic is issued: one shared-memory round trip per item instead of one per tile)
        {
            u32 rr[ITEMS];
#pragma unroll
            for (int j = 0; j < ITEMS; ++j) {
                const u32 i = threadIdx.x + j * THREADS;
                rr[j] = i < n_here ? atomicAdd(&cnt[item[j].digit_top(shift, bits)], 1u) : 0u;
            }
#pragma unroll
            for (int j = 0; j < ITEMS; ++j) {
                const u32 i = threadIdx.x + j * THREADS;
                if (i < n_here) rk[i] = (u16) rr[j];
            }
        }
        __syncthreads();  // counts complete, input buffer free
        u32 b_next = b, n_next = 0;
        u64 first_next = 0;
        if (t + 1 < t1) {
            describe(t + 1, b_next, first_next, n_next);
            prefetch(first_next, n_next);
        }
        const u32 c = threadIdx.x < 256 ? cnt[threadIdx.x] : 0u;
        u32 total;
        const u32 p = kc_block_exclusive_scan<THREADS>(c, &total, sw);
        if (threadIdx.x < 256) {
            loff[threadIdx.x] = p;
            u32 gb = 0;
            if (c) gb = atomicAdd(&C_cnt[((u64) b << bits) + threadIdx.x], c);
            const u32 room = gb < capC ? capC - gb : 0u;
            dbase[threadIdx.x] = (((u64) b << bits) + threadIdx.x) * capC + gb - p;
            qlim[threadIdx.x] = p + (c < room ? c : room);
            cnt[threadIdx.x] = 0;
        }
        __syncthreads();
        {   // staged positions of all items first (rk[] shares the dynamic array with the staging buffers: no load may pass a store)
            u32 qq[ITEMS];
#pragma unroll
            for (int j = 0; j < ITEMS; ++j) {
                const u32 i = threadIdx.x + j * THREADS;
                qq[j] = i < n_here ? (u32) rk[i] : 0u;
            }
#pragma unroll
            for (int j = 0; j < ITEMS; ++j) qq[j] += loff[item[j].digit_top(shift, bits)];
#pragma unroll
            for (int j = 0; j < ITEMS; ++j) {
                const u32 i = threadIdx.x + j * THREADS;
                if (i < n_here) {
                    stage_k[qq[j]] = item[j];
                    stage_p[qq[j]] = pay[j];
                }
            }
        }
        __syncthreads();
#pragma unroll 4
        for (u32 q = threadIdx.x; q < n_here; q += THREADS) {
            const KWord<L> v = stage_k[q];
            const u32 dg = v.digit_top(shift, bits);
            if (q < qlim[dg]) {
                const u64 at = dbase[dg] + q;
                kdst[at] = v;
                pdst[at] = stage_p[q];
            } else {
                over = true;
            }
        }
        b = b_next;
        n_here = n_next;
    }
    if (over) status[0] = 1;
}

// Launches the whole construction on ex.stream and returns without synchronising.
//   cells[0] += distinct k-mers with >= min_freq occurrences, cells[2] = M (k-mer windows), low word of cells[3] = overflow status;
//   flags: zeroed bit array over the n_bytes positions (see kc_kmerset_build).
// Returns false (nothing launched) when the plan does not apply; the caller then uses kc_kmerset_build.
template <int L>
bool kc_kmerset_build_fast(CudaExec &ex, const u8 *seq, u64 n_bytes, int k, bool complements, int min_freq, u32 *flags, u64 *cells,
                           const KsfTuning &tune, KsfPlan *plan_out = nullptr, InputChunks *chunks = nullptr) {
    typedef KsCfg<L> Cfg;
    const KsfPlan pl = kc_ksf_plan(n_bytes, tune);
    if (plan_out) *plan_out = pl;
    if (!pl.ok) return false;
    cudaStream_t st = ex.stream;
    const size_t base_mark = ex.arena->mark();
    const u64 item_bytes = sizeof(KWord<L>) + 4;
    KWord<L> *kb[2];
    u32 *pb[2];
    for (int i = 0; i < 2; ++i) {
        kb[i] = ex.alloc<KWord<L>>(pl.slots[i] ? pl.slots[i] : 1);
        pb[i] = ex.alloc<u32>(pl.slots[i] ? pl.slots[i] : 1);
    }
    u64 n_cnt = 0;
    for (int i = 0; i < pl.n_levels; ++i) n_cnt += 1ULL << pl.cum[i];
    u32 *cnt_all = ex.alloc<u32>(n_cnt);
    ex.fill_bytes(cnt_all, 0, n_cnt * 4);
    u32 *status = reinterpret_cast<u32 *>(cells + 3);
    kc_ull *m_cell = reinterpret_cast<kc_ull *>(cells + 2);

    const int smem0 = Cfg::EX_TILE * ((int) sizeof(KWord<L>) + 4);
    const int tv = tune.tile_variant;
    const int tile1 = (tv == 1 || tv == 4) ? Cfg::TILE / 2 : (tv == 2 ? Cfg::TILE * 3 / 4 : Cfg::TILE);
    const bool pf = tv >= 3;  // input tile double-buffered in shared memory
    const int smem1 = tile1 * ((pf ? 2 : 1) * ((int) sizeof(KWord<L>) + 4) + 2);
    const bool clear_flags = tune.resolve >= 2;  // level 0 writes the valid-window bits, the resolve clears the losers
    const bool counted = min_freq > 1;
    constexpr int CA = KSF_LEAF_CAP;
    const int per_item = (int) sizeof(KWord<L>) + 12 + 4;
    const int smem_r = CA * (per_item + (counted ? 4 : 0));
    static bool attr_done = false;
    static int occ_r[2] = {0, 0}, n_sm = 0;
    if (!attr_done) {
        KC_CUDA(cudaFuncSetAttribute(kc_ksf_scatter0_kernel<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem0));
        KC_CUDA(cudaFuncSetAttribute(kc_ksf_scatter_kernel<L, Cfg::TILE, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     Cfg::TILE * ((int) sizeof(KWord<L>) + 4 + 2)));
        KC_CUDA(cudaFuncSetAttribute(kc_ksf_scatter_kernel<L, Cfg::TILE / 2, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     Cfg::TILE / 2 * ((int) sizeof(KWord<L>) + 4 + 2)));
        KC_CUDA(cudaFuncSetAttribute(kc_ksf_scatter_kernel<L, Cfg::TILE * 3 / 4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     Cfg::TILE * 3 / 4 * ((int) sizeof(KWord<L>) + 4 + 2)));
        KC_CUDA(cudaFuncSetAttribute(kc_ksf_scatter0_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, KsCfg<1>::EX_TILE * 12));
        KC_CUDA(cudaFuncSetAttribute(kc_ksf_scatter_pf_kernel<L, Cfg::TILE, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     Cfg::TILE * (2 * ((int) sizeof(KWord<L>) + 4) + 2)));
        KC_CUDA(cudaFuncSetAttribute(kc_ksf_scatter_pf_kernel<L, Cfg::TILE, 2, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     Cfg::TILE * (2 * ((int) sizeof(KWord<L>) + 4) + 2)));
        KC_CUDA(cudaFuncSetAttribute(kc_ksf_scatter_pf_kernel<L, Cfg::TILE / 2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     Cfg::TILE / 2 * (2 * ((int) sizeof(KWord<L>) + 4) + 2)));
        KC_CUDA(cudaFuncSetAttribute(kc_ks_resolve_hash_kernel<L, CA, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, CA * per_item));
        KC_CUDA(cudaFuncSetAttribute(kc_ks_resolve_hash_kernel<L, CA, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, CA * (per_item + 4)));
        int dev = 0;
        KC_CUDA(cudaGetDevice(&dev));
        KC_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
        KC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_r[0], kc_ks_resolve_hash_kernel<L, CA, false>, 256, CA * per_item));
        KC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_r[1], kc_ks_resolve_hash_kernel<L, CA, true>, 256, CA * (per_item + 4)));
        attr_done = true;
    }

    // ---- level 0 ----
    u32 *cnt_cur = cnt_all;
    {
        const u32 blocks = (u32) kc_div_up(n_bytes, (u64) Cfg::EX_TILE);
        // one launch, or one launch per chunk of an input that is still being copied in (each waits for its own chunk only)
        const int n_parts = chunks && chunks->n > 1 && !chunks->waited ? chunks->n : 1;
        if (chunks && n_parts == 1) chunks->wait_all(st);
        for (int c = 0; c < n_parts; ++c) {
            u32 t0 = 0, t1 = blocks;
            if (n_parts > 1) {
                KC_CUDA(cudaStreamWaitEvent(st, chunks->ev[c], 0));
                t0 = (u32) std::min<u64>(blocks, (u64) c * (chunks->chunk_bytes / Cfg::EX_TILE));
                t1 = c == n_parts - 1 ? blocks : (u32) std::min<u64>(blocks, (u64) (c + 1) * (chunks->chunk_bytes / Cfg::EX_TILE));
            }
            if (t1 <= t0) continue;
            const u64 part_bytes = std::min<u64>(n_bytes, (u64) t1 * Cfg::EX_TILE) - (u64) t0 * Cfg::EX_TILE;
            CudaExec::Scope sc(ex, KP_KS_SCATTER0, part_bytes + part_bytes * item_bytes);
            bool split = false;
            if constexpr (L == 1) {
                if (tune.split0 && !clear_flags) {
                    split = true;
                    kc_ksf_scatter0_split_kernel<<<t1 - t0, 512, smem0, st>>>(seq, n_bytes, k, complements ? 1 : 0, 64 * L - pl.cum[0], pl.bits[0], cnt_cur,
                                                                              (u32) pl.cap[0], kb[0], pb[0], status, t0);
                }
            }
            if (!split)
                kc_ksf_scatter0_kernel<L><<<t1 - t0, Cfg::EX_THREADS, smem0, st>>>(seq, n_bytes, k, complements ? 1 : 0, 64 * L - pl.cum[0], pl.bits[0], cnt_cur,
                                                                                 (u32) pl.cap[0], kb[0], pb[0], status, t0, clear_flags ? flags : nullptr,
                                                                                 (u32) (kc_div_up(n_bytes, (u64) 32) + 1));  // = kc_runs_flag_words(n_bytes)
            ++ex.launches;
            KC_CUDA(cudaGetLastError());
        }
        if (chunks) chunks->waited = true;
    }
    // ---- levels >= 1 ----
    const u32 max_ctas = tune.max_ctas > 0 ? (u32) tune.max_ctas : 148 * 8;
    for (int lv = 1; lv < pl.n_levels; ++lv) {
        const u32 nP = (u32) (1ULL << pl.cum[lv - 1]);
        const size_t mark = ex.arena->mark();
        u32 *P_size = ex.alloc<u32>(nP);
        u32 *tile_prefix = ex.alloc<u32>((u64) nP + 1);
        kc_ksf_prep_kernel<<<(unsigned) kc_div_up((u64) nP + 1, 256), 256, 0, st>>>(cnt_cur, nP, (u32) pl.cap[lv - 1], (u32) tile1, P_size, tile_prefix,
                                                                                  status, lv == 1 ? m_cell : nullptr);
        ++ex.launches;
        ex.exclusive_scan_nosync(tile_prefix, tile_prefix, (u64) nP + 1);
        u32 *cnt_next = cnt_cur + nP;
        const u64 tiles_ub = n_bytes / tile1 + nP + 1;
        const u32 tiles_per_cta = (u32) kc_div_up(tiles_ub, max_ctas);
        const u32 ctas = (u32) kc_div_up(tiles_ub, tiles_per_cta);
        {
            CudaExec::Scope sc(ex, KP_SORT_SCATTER, 2 * n_bytes * item_bytes);
            if (tv == 3) kc_ksf_scatter_pf_kernel<L, Cfg::TILE, 2><<<ctas, 256, smem1, st>>>(kb[(lv - 1) & 1], pb[(lv - 1) & 1], kb[lv & 1], pb[lv & 1], P_size, tile_prefix, nP,
                                                               pl.cap[lv - 1], tiles_per_cta, 64 * L - pl.cum[lv], pl.bits[lv], cnt_next, (u32) pl.cap[lv],
                                                               status);
            else if (tv == 5) kc_ksf_scatter_pf_kernel<L, Cfg::TILE, 2, 512><<<ctas, 512, smem1, st>>>(kb[(lv - 1) & 1], pb[(lv - 1) & 1], kb[lv & 1], pb[lv & 1], P_size, tile_prefix, nP,
                                                               pl.cap[lv - 1], tiles_per_cta, 64 * L - pl.cum[lv], pl.bits[lv], cnt_next, (u32) pl.cap[lv],
                                                               status);
            else if (tv == 4) kc_ksf_scatter_pf_kernel<L, Cfg::TILE / 2, 4><<<ctas, 256, smem1, st>>>(kb[(lv - 1) & 1], pb[(lv - 1) & 1], kb[lv & 1], pb[lv & 1], P_size, tile_prefix, nP,
                                                               pl.cap[lv - 1], tiles_per_cta, 64 * L - pl.cum[lv], pl.bits[lv], cnt_next, (u32) pl.cap[lv],
                                                               status);
            else if (tv == 1) kc_ksf_scatter_kernel<L, Cfg::TILE / 2, 5><<<ctas, 256, smem1, st>>>(kb[(lv - 1) & 1], pb[(lv - 1) & 1], kb[lv & 1], pb[lv & 1], P_size, tile_prefix, nP,
                                                               pl.cap[lv - 1], tiles_per_cta, 64 * L - pl.cum[lv], pl.bits[lv], cnt_next, (u32) pl.cap[lv],
                                                               status);
            else if (tv == 2) kc_ksf_scatter_kernel<L, Cfg::TILE * 3 / 4, 4><<<ctas, 256, smem1, st>>>(kb[(lv - 1) & 1], pb[(lv - 1) & 1], kb[lv & 1], pb[lv & 1], P_size, tile_prefix, nP,
                                                               pl.cap[lv - 1], tiles_per_cta, 64 * L - pl.cum[lv], pl.bits[lv], cnt_next, (u32) pl.cap[lv],
                                                               status);
            else kc_ksf_scatter_kernel<L, Cfg::TILE, 3><<<ctas, 256, smem1, st>>>(kb[(lv - 1) & 1], pb[(lv - 1) & 1], kb[lv & 1], pb[lv & 1], P_size, tile_prefix, nP,
                                                               pl.cap[lv - 1], tiles_per_cta, 64 * L - pl.cum[lv], pl.bits[lv], cnt_next, (u32) pl.cap[lv],
                                                               status);
            ++ex.launches;
            KC_CUDA(cudaGetLastError());
        }
        cnt_cur = cnt_next;
        ex.arena->release(mark);  // stream order keeps P_size / tile_prefix alive until the scatter has run
    }
    // ---- leaves -> hash resolve ----
    const int last = pl.n_levels - 1;
    if (pl.n_leaf >= 0xFFFFFFFFULL) KC_THROW(KC_ERR_TOO_LARGE, "too many leaf buckets");
    const u32 n_small = (u32) pl.n_leaf;
    kc_ull *n_unique = reinterpret_cast<kc_ull *>(cells);
    if (pl.n_levels == 1) {  // M has not been added up by a prep kernel
        const u32 *cc = cnt_cur;
        const u64 nl = pl.n_leaf;
        ex.for_each(nl, [=] __device__(u64 i) {
            const u32 c = cc[i] < KSF_LEAF_CAP ? cc[i] : KSF_LEAF_CAP;
            if (c) atomicAdd(m_cell, (kc_ull) c);
        });
    }
    if (tune.resolve >= 6 && !counted) {
        const int smem4 = (int) KSF_LEAF_CAP * (2 * (8 * L + 4) + 16);
        static bool attr4_done = false;
        static int occ4[2] = {0, 0};
        if (!attr4_done) {
            KC_CUDA(cudaFuncSetAttribute(kc_ksf_resolve2_kernel<L, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem4));
            KC_CUDA(cudaFuncSetAttribute(kc_ksf_resolve2_kernel<L, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem4));
            KC_CUDA(cudaFuncSetAttribute(kc_ksf_resolve2_kernel<L, 256, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem4));
            KC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ4[0], kc_ksf_resolve2_kernel<L, 256>, 256, smem4));
            KC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ4[1], kc_ksf_resolve2_kernel<L, 512>, 512, smem4));
            attr4_done = true;
        }
        const int wide = tune.resolve == 7 ? 1 : 0;
        const u32 fit = (u32) (n_sm * (occ4[wide] > 0 ? occ4[wide] : 1));
        const u32 grid = n_small < fit ? n_small : fit;
        CudaExec::Scope sc(ex, KP_KS_RESOLVE, n_bytes * item_bytes);
        if (tune.resolve == 8) kc_ksf_resolve2_kernel<L, 256, true><<<grid, 256, smem4, st>>>(kb[last & 1], pb[last & 1], cnt_cur, n_small, flags, n_unique, status);
        else if (wide) kc_ksf_resolve2_kernel<L, 512><<<grid, 512, smem4, st>>>(kb[last & 1], pb[last & 1], cnt_cur, n_small, flags, n_unique, status);
        else kc_ksf_resolve2_kernel<L, 256><<<grid, 256, smem4, st>>>(kb[last & 1], pb[last & 1], cnt_cur, n_small, flags, n_unique, status);
        ++ex.launches;
        KC_CUDA(cudaGetLastError());
    } else if (tune.resolve >= 3 && !counted) {
        const int stages = tune.resolve == 3 ? 2 : (tune.resolve == 4 ? 3 : 4);
        // per ring stage: keys + positions of one leaf; tables: T2 x2 (u32), T1 x2 (two u16 slots per item)
        const int smem3 = (int) KSF_LEAF_CAP * (stages * (8 * L + 4) + 8 + 8);
        static bool attr3_done = false;
        static int occ3[3] = {0, 0, 0};
        if (!attr3_done) {
            const int per = 8 * L + 4;
            KC_CUDA(cudaFuncSetAttribute(kc_ksf_resolve1_kernel<L, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) KSF_LEAF_CAP * (2 * per + 16)));
            KC_CUDA(cudaFuncSetAttribute(kc_ksf_resolve1_kernel<L, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) KSF_LEAF_CAP * (3 * per + 16)));
            KC_CUDA(cudaFuncSetAttribute(kc_ksf_resolve1_kernel<L, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) KSF_LEAF_CAP * (4 * per + 16)));
            KC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ3[0], kc_ksf_resolve1_kernel<L, 2>, 256, (int) KSF_LEAF_CAP * (2 * per + 16)));
            KC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ3[1], kc_ksf_resolve1_kernel<L, 3>, 256, (int) KSF_LEAF_CAP * (3 * per + 16)));
            KC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ3[2], kc_ksf_resolve1_kernel<L, 4>, 256, (int) KSF_LEAF_CAP * (4 * per + 16)));
            attr3_done = true;
        }
        const int occ = occ3[stages - 2];
        const u32 fit = (u32) (n_sm * (occ > 0 ? occ : 1));
        const u32 grid = n_small < fit ? n_small : fit;
        CudaExec::Scope sc(ex, KP_KS_RESOLVE, n_bytes * item_bytes);
        if (stages == 2) kc_ksf_resolve1_kernel<L, 2><<<grid, 256, smem3, st>>>(kb[last & 1], pb[last & 1], cnt_cur, n_small, flags, n_unique, status);
        else if (stages == 3) kc_ksf_resolve1_kernel<L, 3><<<grid, 256, smem3, st>>>(kb[last & 1], pb[last & 1], cnt_cur, n_small, flags, n_unique, status);
        else kc_ksf_resolve1_kernel<L, 4><<<grid, 256, smem3, st>>>(kb[last & 1], pb[last & 1], cnt_cur, n_small, flags, n_unique, status);
        ++ex.launches;
        KC_CUDA(cudaGetLastError());
    } else if (tune.resolve >= 1) {
        const int smem2 = (int) KSF_LEAF_CAP * (16 * L + 20 + (counted ? 4 : 0));
        static bool attr2_done = false;
        static int occ2[2] = {0, 0};
        if (!attr2_done) {
            KC_CUDA(cudaFuncSetAttribute(kc_ksf_resolve_kernel<L, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) KSF_LEAF_CAP * (16 * L + 20)));
            KC_CUDA(cudaFuncSetAttribute(kc_ksf_resolve_kernel<L, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) KSF_LEAF_CAP * (16 * L + 24)));
            KC_CUDA(cudaFuncSetAttribute(kc_ksf_resolve_kernel<L, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) KSF_LEAF_CAP * (16 * L + 20)));
            KC_CUDA(cudaFuncSetAttribute(kc_ksf_resolve_kernel<L, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) KSF_LEAF_CAP * (16 * L + 24)));
            KC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2[0], kc_ksf_resolve_kernel<L, false>, 256, (int) KSF_LEAF_CAP * (16 * L + 20)));
            KC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2[1], kc_ksf_resolve_kernel<L, true>, 256, (int) KSF_LEAF_CAP * (16 * L + 24)));
            attr2_done = true;
        }
        const u32 fit = (u32) (n_sm * (occ2[counted] > 0 ? occ2[counted] : 1));
        const u32 grid = n_small < fit ? n_small : fit;
        CudaExec::Scope sc(ex, KP_KS_RESOLVE, n_bytes * item_bytes);
        if (counted && clear_flags)
            kc_ksf_resolve_kernel<L, true, true><<<grid, 256, smem2, st>>>(kb[last & 1], pb[last & 1], cnt_cur, n_small, flags, (u32) min_freq, n_unique, status);
        else if (counted)
            kc_ksf_resolve_kernel<L, true><<<grid, 256, smem2, st>>>(kb[last & 1], pb[last & 1], cnt_cur, n_small, flags, (u32) min_freq, n_unique, status);
        else if (clear_flags)
            kc_ksf_resolve_kernel<L, false, true><<<grid, 256, smem2, st>>>(kb[last & 1], pb[last & 1], cnt_cur, n_small, flags, (u32) min_freq, n_unique, status);
        else
            kc_ksf_resolve_kernel<L, false><<<grid, 256, smem2, st>>>(kb[last & 1], pb[last & 1], cnt_cur, n_small, flags, (u32) min_freq, n_unique, status);
        ++ex.launches;
        KC_CUDA(cudaGetLastError());
    } else {
        SortBucket *small = ex.alloc<SortBucket>(pl.n_leaf);
        kc_ksf_leaf_kernel<<<(unsigned) kc_div_up(pl.n_leaf, 256), 256, 0, st>>>(cnt_cur, pl.n_leaf, KSF_LEAF_CAP, (u8) (last & 1), small, status);
        ++ex.launches;
        const u32 fit = (u32) (n_sm * (occ_r[counted] > 0 ? occ_r[counted] : 1));
        const u32 grid = n_small < fit ? n_small : fit;
        CudaExec::Scope sc(ex, KP_KS_RESOLVE, n_bytes * item_bytes);
        if (counted)
            kc_ks_resolve_hash_kernel<L, CA, true><<<grid, 256, smem_r, st>>>(kb[0], kb[1], pb[0], pb[1], small, n_small, 0u, (u32) CA, flags,
                                                                              (u32) min_freq, n_unique);
        else
            kc_ks_resolve_hash_kernel<L, CA, false><<<grid, 256, smem_r, st>>>(kb[0], kb[1], pb[0], pb[1], small, n_small, 0u, (u32) CA, flags,
                                                                               (u32) min_freq, n_unique);
        ++ex.launches;
        KC_CUDA(cudaGetLastError());
    }
    ex.arena->release(base_mark);
    return true;
}

// Multi-GPU owner side (kc_p2p_resolve): the rank's level-0 buckets arrive complete and back to back in the receive buffer
// (sh->pre_off / pre_size).  The same fixed-slot levels + leaf resolve as above take over from there, instead of the
// histogram-based levels of kmerset.cuh: no counting passes and no host read-backs between the kernels.  Returns false —
// with NOTHING written to `flags` — when the plan does not apply or a slot overflowed (checked on the device before the
// resolve is launched); the caller then runs the exact construction.
template <int L>
bool kc_kmerset_resolve_fast(CudaExec &ex, int k, int min_freq, u32 *flags, const KsShard *sh, const KsfTuning &tune, u64 *n_kept_out) {
    typedef KsCfg<L> Cfg;
    (void) k;
    const u32 nP0 = sh->n_pre;
    const u64 total = sh->n_items;
    if (!tune.enabled || nP0 == 0 || total < tune.min_items) return false;
    u64 max0 = 0;
    for (u32 i = 0; i < nP0; ++i) max0 = sh->pre_size[i] > max0 ? sh->pre_size[i] : max0;
    if (max0 >= 0xFFFFFFF0ULL) return false;
    int sub_bits = 0;
    while (sub_bits < 32 && (max0 >> sub_bits) > tune.leaf_target) ++sub_bits;
    if (sub_bits == 0) return false;
    const int levels = (sub_bits + 7) / 8;
    if (levels > KSF_MAX_LEVELS - 1) return false;
    int bits[KSF_MAX_LEVELS], cum[KSF_MAX_LEVELS];
    u64 cap[KSF_MAX_LEVELS], slots[2] = {1, 1};
    int c = 0;
    for (int i = 0; i < levels; ++i) {
        bits[i] = sub_bits / levels + (i < sub_bits % levels ? 1 : 0);
        c += bits[i];
        cum[i] = c;
        const double mean = (double) max0 / (double) (1ULL << c);
        u64 cp = (u64) std::ceil(mean + tune.sigmas * std::sqrt(mean) + (tune.sigmas > 0 ? 0.02 * mean + 64.0 : 0.0));
        cp = (cp + 31) / 32 * 32;
        if (i == levels - 1) cp = KSF_LEAF_CAP;
        cap[i] = cp;
        const u64 need = ((u64) nP0 << c) * cp;
        if (need > slots[i & 1]) slots[i & 1] = need;
    }
    const u64 n_leaf = (u64) nP0 << cum[levels - 1];
    if (n_leaf >= 0xFFFFFFFFULL) return false;
    cudaStream_t st = ex.stream;
    const size_t base_mark = ex.arena->mark();
    u64 n_cnt = 0;
    for (int i = 0; i < levels; ++i) n_cnt += (u64) nP0 << cum[i];
    {   // the slots must fit what is left of the arena (the caller falls back to the exact construction otherwise)
        const u64 need = (slots[0] + slots[1]) * (sizeof(KWord<L>) + 4) + n_cnt * 4 + ((u64) nP0 << cum[levels - 1]) * 8 + (1u << 20);
        if (ex.arena->off + need > ex.arena->top) return false;
    }
    KWord<L> *kb[2];
    u32 *pb[2];
    for (int i = 0; i < 2; ++i) {
        kb[i] = ex.alloc<KWord<L>>(slots[i]);
        pb[i] = ex.alloc<u32>(slots[i]);
    }
    u32 *cnt_all = ex.alloc<u32>(n_cnt);
    ex.fill_bytes(cnt_all, 0, n_cnt * 4);
    u64 *d_off = ex.alloc<u64>(nP0);
    u32 *d_size = ex.alloc<u32>(nP0);
    u64 *cells = ex.alloc<u64>(2);  // [0] kept, [1] low word = overflow status
    ex.fill_bytes(cells, 0, 16);
    u32 *status = reinterpret_cast<u32 *>(cells + 1);
    kc_ull *n_unique = reinterpret_cast<kc_ull *>(cells);
    {
        std::vector<u32> hs(nP0);
        for (u32 i = 0; i < nP0; ++i) hs[i] = (u32) sh->pre_size[i];
        KC_CUDA(cudaMemcpyAsync(d_off, sh->pre_off, (size_t) nP0 * 8, cudaMemcpyHostToDevice, st));
        KC_CUDA(cudaMemcpyAsync(d_size, hs.data(), (size_t) nP0 * 4, cudaMemcpyHostToDevice, st));
        KC_CUDA(cudaStreamSynchronize(st));  // hs goes out of scope
    }
    const int tile1 = Cfg::TILE;
    const int smem1 = tile1 * ((int) sizeof(KWord<L>) + 4 + 2);
    static bool attr_done = false;
    if (!attr_done) {
        KC_CUDA(cudaFuncSetAttribute(kc_ksf_scatter_kernel<L, Cfg::TILE, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem1));
        KC_CUDA(cudaFuncSetAttribute(kc_ksf_resolve_kernel<L, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) KSF_LEAF_CAP * (16 * L + 20)));
        KC_CUDA(cudaFuncSetAttribute(kc_ksf_resolve_kernel<L, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) KSF_LEAF_CAP * (16 * L + 24)));
        attr_done = true;
    }
    const u64 item_bytes = sizeof(KWord<L>) + 4;
    const u32 max_ctas = 148 * 8;
    const KWord<L> *ksrc = reinterpret_cast<const KWord<L> *>(sh->keys);
    const u32 *psrc = sh->pos;
    const u32 *cnt_par = d_size;
    u32 *cnt_cur = cnt_all;
    for (int lv = 0; lv < levels; ++lv) {
        const u32 nP = lv == 0 ? nP0 : (u32) ((u64) nP0 << cum[lv - 1]);
        const size_t mark = ex.arena->mark();
        u32 *P_size = ex.alloc<u32>(nP);
        u32 *tile_prefix = ex.alloc<u32>((u64) nP + 1);
        const u32 capP = lv == 0 ? 0xFFFFFFFFu : (u32) cap[lv - 1];
        kc_ksf_prep_kernel<<<(unsigned) kc_div_up((u64) nP + 1, 256), 256, 0, st>>>(cnt_par, nP, capP, (u32) tile1, P_size, tile_prefix, status, nullptr);
        ++ex.launches;
        ex.exclusive_scan_nosync(tile_prefix, tile_prefix, (u64) nP + 1);
        const u64 tiles_ub = total / tile1 + nP + 1;
        const u32 tiles_per_cta = (u32) kc_div_up(tiles_ub, max_ctas);
        const u32 ctas = (u32) kc_div_up(tiles_ub, tiles_per_cta);
        {
            CudaExec::Scope sc(ex, KP_SORT_SCATTER, 2 * total * item_bytes);
            kc_ksf_scatter_kernel<L, Cfg::TILE, 3><<<ctas, 256, smem1, st>>>(ksrc, psrc, kb[lv & 1], pb[lv & 1], P_size, tile_prefix, nP, (u64) capP, tiles_per_cta,
                                                                             64 * L - Cfg::D0 - cum[lv], bits[lv], cnt_cur, (u32) cap[lv], status,
                                                                             lv == 0 ? d_off : nullptr);
            ++ex.launches;
            KC_CUDA(cudaGetLastError());
        }
        ksrc = kb[lv & 1];
        psrc = pb[lv & 1];
        cnt_par = cnt_cur;
        cnt_cur += (u64) nP0 << cum[lv];
        ex.arena->release(mark);
    }
    {   // no leaf may exceed the resolve capacity: decided BEFORE a single flag bit is written
        const u32 *cc = cnt_par;
        ex.for_each(n_leaf, [=] __device__(u64 i) {
            if (cc[i] > KSF_LEAF_CAP) status[0] = 1;
        });
    }
    if (ex.read(status)) {
        ex.arena->release(base_mark);
        return false;
    }
    const bool counted = min_freq > 1;
    const int smem2 = (int) KSF_LEAF_CAP * (16 * L + 20 + (counted ? 4 : 0));
    int occ = 0, n_sm = 0, dev = 0;
    KC_CUDA(cudaGetDevice(&dev));
    KC_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    if (counted) KC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kc_ksf_resolve_kernel<L, true>, 256, smem2));
    else KC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kc_ksf_resolve_kernel<L, false>, 256, smem2));
    const u32 fit = (u32) (n_sm * (occ > 0 ? occ : 1));
    const u32 grid = (u32) n_leaf < fit ? (u32) n_leaf : fit;
    {
        CudaExec::Scope sc(ex, KP_KS_RESOLVE, total * item_bytes);
        const int last = levels - 1;
        if (counted)
            kc_ksf_resolve_kernel<L, true><<<grid, 256, smem2, st>>>(kb[last & 1], pb[last & 1], cnt_par, (u32) n_leaf, flags, (u32) min_freq, n_unique, status);
        else
            kc_ksf_resolve_kernel<L, false><<<grid, 256, smem2, st>>>(kb[last & 1], pb[last & 1], cnt_par, (u32) n_leaf, flags, (u32) min_freq, n_unique, status);
        ++ex.launches;
        KC_CUDA(cudaGetLastError());
    }
    *n_kept_out = ex.read(cells);
    ex.arena->release(base_mark);
    return true;
}

#endif  // __CUDACC__

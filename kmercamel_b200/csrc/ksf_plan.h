// Plan of the histogram-free (fixed-slot) k-mer set construction of kmerset_fast.cuh: how many MSD partition levels, how many
// key bits each consumes, and the slot size of every level, all from the input size alone.  Pure host C++ (no CUDA), so that
// tests/host_emul can check the invariants the kernels rely on:
//   * a level's digit is at most 8 bits wide (the kernels keep 256 counters per tile);
//   * every slot size is a multiple of 32 items, so slots stay 128-byte aligned for 4-byte positions and 256-byte aligned for
//     8-byte keys — the cp.async (16-byte) tile and leaf copies of kc_ksf_scatter_pf_kernel / kc_ksf_resolve*_kernel need that;
//   * all partition digits together stay inside the top limb of the scrambled word (KWord::digit_top);
//   * a leaf slot holds KSF_LEAF_CAP items, the capacity of the shared-memory resolve.
#pragma once
#include <cmath>
#include <cstdint>

static const uint32_t KSF_LEAF_CAP = 1024;  // capacity of kc_ks_resolve_hash_kernel<L, 1024, *>
static const int KSF_MAX_LEVELS = 4;

struct KsfTuning {
    bool enabled = true;
    uint32_t leaf_target = 768;   // mean leaf size aimed at (the leaf slot holds KSF_LEAF_CAP)
    double sigmas = 8.0;     // slot = mean + sigmas * sqrt(mean) (+ 2 %) items; 0 forces overflows (tests)
    uint64_t min_items = 1u << 16;  // smaller inputs take the exact path (nothing to gain)
    int resolve = 6;         // 1 = kc_ksf_resolve_kernel (cp.async, two barriers), 0 = kc_ks_resolve_hash_kernel over a bucket list,
                             // 2 = as 1 with clear-the-losers flags (level 0 writes the valid-window bits, a duplicate clears one),
                             // 3 = kc_ksf_resolve1_kernel: clear-the-losers + double-buffered tables, ONE barrier per leaf (-z 1 only),
                             // 4 / 5 = the same with 3 / 4 staging buffers in the ring (two / three leaves in flight per CTA),
                             // 6 / 7 = kc_ksf_resolve2_kernel: as 3 with a thread's items staged in registers, 256 / 512 threads,
                             // 8 = as 6 with 8-byte copy units (every thread the same number of items) and the leaf size loaded a leaf ahead
    int tile_variant = 5;    // level >= 1 scatter: 0 = KsCfg<L>::TILE items per tile, 3 CTAs/SM; 1 = half tiles, 5 CTAs/SM; 2 = 3/4 tiles, 4 CTAs/SM;
                             // 3 / 4 = kc_ksf_scatter_pf_kernel (next tile streams in with cp.async) with full / half tiles,
                             // 5 = full tiles and 512-thread CTAs (2 per SM: 32 instead of 16 warps)
    int max_ctas = 148 * 16; // level >= 1 scatter grid: at most this many CTAs, each walking a run of consecutive tiles (0 = 148 * 8)
    int split0 = 0;          // level 0 (L = 1): two threads per 32-base strip (512-thread CTAs); measured slower (0.327 vs 0.315 ms), kept as an option
};

struct KsfPlan {
    bool ok = false;
    int n_levels = 0;
    int bits[KSF_MAX_LEVELS];   // digit width of level i
    int cum[KSF_MAX_LEVELS];    // key bits consumed after level i
    uint64_t cap[KSF_MAX_LEVELS];    // slot size (items) of one bucket produced by level i
    uint64_t slots[2] = {0, 0};      // items the ping (even levels) / pong (odd levels) buffers must hold
    uint64_t n_leaf = 0;
};

inline KsfPlan kc_ksf_plan(uint64_t m_upper, const KsfTuning &t) {
    KsfPlan p;
    if (!t.enabled || m_upper < t.min_items) return p;
    int total_bits = 1;
    while (total_bits < 40 && (m_upper >> total_bits) > t.leaf_target) ++total_bits;
    const int levels = (total_bits + 7) / 8;
    if (levels > KSF_MAX_LEVELS) return p;
    p.n_levels = levels;
    int cum = 0;
    for (int i = 0; i < levels; ++i) {
        p.bits[i] = total_bits / levels + (i < total_bits % levels ? 1 : 0);
        cum += p.bits[i];
        p.cum[i] = cum;
        const double mean = (double) m_upper / (double) (1ULL << cum);
        uint64_t cap = (uint64_t) std::ceil(mean + t.sigmas * std::sqrt(mean) + (t.sigmas > 0 ? 0.02 * mean + 64.0 : 0.0));
        cap = (cap + 31) / 32 * 32;
        if (i == levels - 1) cap = KSF_LEAF_CAP;
        if (cap >= 0xFFFFFFFFULL) return p;
        p.cap[i] = cap;
        const uint64_t need = (1ULL << cum) * cap;
        if (need > p.slots[i & 1]) p.slots[i & 1] = need;
    }
    p.n_leaf = 1ULL << cum;
    p.ok = true;
    return p;
}


// TEST INFRASTRUCTURE ONLY — never linked into or called by the product.
//
// Thin driver around the UNMODIFIED reference headers (found through -I/root/reference/src at build
// time; nothing is copied into this repo).  The `kmercamel` CLI of the reference only exposes the
// final .msfa, so parity tests for the intermediate stages need these dumps:
//
//   ref_harness kmers  <fasta> <k> <complements:0|1> <out.bin>
//       canonical k-mer -> (occurrences-1 saturated to 255) exactly as AddKMersWithFrequencies
//       (reference src/parser.h:53-85) builds it; written sorted by key as
//       u64 n, u32 limbs, then n * (limbs * u64 little-endian limbs), then n * u8.
//   ref_harness path   <fasta> <k> <complements:0|1> <lower_bound:0|1>
//       records are nodes in file order (simplitigs_from_fasta, src/simplitigs.h:89-103);
//       prints edgeFrom / overlaps of OverlapHamiltonianPath (src/global.h:43-133), -1 / 255 = none.
//   ref_harness sparsepath <fasta> <k> <complements:0|1> <lower_bound:0|1>
//       every record must be exactly k long; OverlapHamiltonianPathSparse (src/global_sparse.h:42-132).
//   ref_harness full   <fasta> <k> <complements:0|1> <min_frequency> <out.ms|->
//       the k-mer stage of kmercamel<>() (src/main.cpp:146-162) followed — unless the output path is "-" — by
//       get_simplitigs + Global / GlobalSparse exactly as src/main.cpp:169-186 chains them, in ONE process, so that
//       a full-size input is hashed once.  Prints one JSON line: n_kmers, an order-independent digest of the kept
//       (k-mer, min(occurrences,256)) pairs (see kmer_digest below), n_simplitigs, sparse switch taken, superstring
//       length and ones, stage seconds.  The superstring line goes to <out.ms>.
//
// Include order: ac/kmers_ac.h must come first.
#include "ac/kmers_ac.h"
#include "parser.h"
#include "simplitigs.h"
#include "global.h"
#include "global_sparse.h"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <fstream>
#include <string>
#include <vector>

template <typename kmer_t> static void to_limbs(kmer_t v, uint64_t *out, int limbs);
template <> void to_limbs<kmer64_t>(kmer64_t v, uint64_t *out, int) { out[0] = v; }
template <> void to_limbs<kmer128_t>(kmer128_t v, uint64_t *out, int) {
    out[0] = (uint64_t) v;
    out[1] = (uint64_t) (v >> 64);
}
template <> void to_limbs<kmer256_t>(kmer256_t v, uint64_t *out, int) {
    out[0] = (uint64_t) v.lower();
    out[1] = (uint64_t) (v.lower() >> 64);
    out[2] = (uint64_t) v.upper();
    out[3] = (uint64_t) (v.upper() >> 64);
}

template <typename kmer_t, typename wrapper_t>
static int dump_kmers(wrapper_t wrapper, std::string path, int k, bool complements, const char *out_path) {
    auto *freq = wrapper.kh_init_freq_map();
    gzFile fp = OpenFile(path);
    kseq_t *seq = kseq_init(fp);
    while (kseq_read(seq) >= 0)
        AddKMersWithFrequencies(freq, wrapper, kmer_t(0), seq->seq.l, seq->seq.s, k, complements);
    kseq_destroy(seq);
    gzclose(fp);
    std::vector<std::pair<kmer_t, uint8_t>> all;
    all.reserve(kh_size(freq));
    for (auto i = kh_begin(freq); i != kh_end(freq); ++i)
        if (kh_exist(freq, i)) all.emplace_back(kh_key(freq, i), kh_val(freq, i));
    std::sort(all.begin(), all.end(), [](const auto &a, const auto &b) { return a.first < b.first; });
    const uint32_t limbs = sizeof(kmer_t) / 8;
    FILE *f = std::fopen(out_path, "wb");
    if (!f) return 2;
    uint64_t n = all.size();
    std::fwrite(&n, 8, 1, f);
    std::fwrite(&limbs, 4, 1, f);
    std::vector<uint64_t> buf(n * limbs);
    for (uint64_t i = 0; i < n; ++i) to_limbs<kmer_t>(all[i].first, &buf[i * limbs], limbs);
    std::fwrite(buf.data(), 8, buf.size(), f);
    std::vector<uint8_t> vals(n);
    for (uint64_t i = 0; i < n; ++i) vals[i] = all[i].second;
    std::fwrite(vals.data(), 1, n, f);
    std::fclose(f);
    return 0;
}

// Order-independent digest of a k-mer set with counts: h(key) folds the little-endian limbs through the splitmix64
// finaliser; the digest is (n, sum h, xor h, sum h * c) mod 2^64 with c = min(occurrences, 256).  The product computes the
// same four numbers on the device (kc_kmer_digest), so sets of 10^9 keys are compared without moving or sorting them.
static inline uint64_t mix64(uint64_t x) {
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ULL;
    x ^= x >> 27; x *= 0x94d049bb133111ebULL;
    x ^= x >> 31;
    return x;
}
static inline uint64_t limb_hash(const uint64_t *l, int limbs) {
    uint64_t h = mix64(l[0] + 0x9e3779b97f4a7c15ULL);
    for (int i = 1; i < limbs; ++i) h = mix64(h ^ (l[i] + 0x9e3779b97f4a7c15ULL * (uint64_t) (i + 1)));
    return h;
}
struct kmer_digest {
    uint64_t n = 0, sum = 0, x = 0, wsum = 0;
    void add(const uint64_t *l, int limbs, uint64_t c) {
        uint64_t h = limb_hash(l, limbs);
        ++n; sum += h; x ^= h; wsum += h * c;
    }
};

static double now_s() {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

template <typename kmer_t, typename wrapper_t>
static int run_full(wrapper_t wrapper, std::string path, int k, bool complements, int z, const char *out_path) {
    const int limbs = sizeof(kmer_t) / 8;
    double t0 = now_s();
    kmer_digest dg;
    auto *kMers = wrapper.kh_init_set();
    uint64_t lb[4];
    if (z == 1) {                                   // src/main.cpp:149-150
        ReadKMers(kMers, wrapper, kmer_t(0), path, k, complements);
        for (auto i = kh_begin(kMers); i != kh_end(kMers); ++i)
            if (kh_exist(kMers, i)) { to_limbs<kmer_t>(kh_key(kMers, i), lb, limbs); dg.add(lb, limbs, 1); }
    } else {                                        // ReadKMersFiltered, src/parser.h:122-141, with the digest taken from the map
        gzFile fp = OpenFile(path);
        kseq_t *seq = kseq_init(fp);
        auto freq = wrapper.kh_init_freq_map();
        while (kseq_read(seq) >= 0) AddKMersWithFrequencies(freq, wrapper, kmer_t(0), seq->seq.l, seq->seq.s, k, complements);
        for (auto i = kh_begin(freq); i != kh_end(freq); ++i)
            if (kh_exist(freq, i) && ((uint16_t) kh_val(freq, i)) + 1 >= z) {
                to_limbs<kmer_t>(kh_key(freq, i), lb, limbs);
                dg.add(lb, limbs, (uint64_t) kh_val(freq, i) + 1);
            }
        auto vec = kMersToVecFiltered(freq, kmer_t(0), (uint16_t) z);
        wrapper.kh_destroy_freq_map(freq);
        kMersFromVec(kMers, wrapper, vec);
        kseq_destroy(seq);
        gzclose(fp);
    }
    size_t kmer_count = kh_size(kMers);
    double t1 = now_s();
    std::printf("{\"n_kmers\": %llu, \"digest\": [%llu, %llu, %llu, %llu], \"kmers_s\": %.1f", (unsigned long long) kmer_count,
                (unsigned long long) dg.n, (unsigned long long) dg.sum, (unsigned long long) dg.x, (unsigned long long) dg.wsum, t1 - t0);
    std::fflush(stdout);
    if (std::strcmp(out_path, "-") != 0 && kmer_count) {
        std::ofstream of(out_path);
        auto simplitigs = get_simplitigs(kMers, wrapper, kmer_t(0), k, complements);   // src/main.cpp:169-171
        wrapper.kh_destroy_set(kMers);
        double t2 = now_s();
        bool sparse = simplitigs.size() * 5 >= kmer_count;                             // src/main.cpp:94,175
        size_t n_simplitigs = simplitigs.size();
        if (sparse) {
            auto vec = simplitigs_to_kmer_vec(kmer_t(0), simplitigs, k, kmer_count);
            PartialPreSort(vec, k);
            GlobalSparse(wrapper, vec, of, nullptr, k, complements);
        } else {
            Global(wrapper, kmer_t(0), simplitigs, of, nullptr, k, complements);
        }
        of << std::endl;
        of.close();
        double t3 = now_s();
        std::printf(", \"n_simplitigs\": %llu, \"sparse\": %d, \"simplitigs_s\": %.1f, \"global_s\": %.1f", (unsigned long long) n_simplitigs,
                    (int) sparse, t2 - t1, t3 - t2);
    }
    std::printf(", \"total_s\": %.1f}\n", now_s() - t0);
    return 0;
}

static void print_path(const overlapPath &p) {
    for (size_t i = 0; i < p.first.size(); ++i)
        std::printf("%lld %d\n", p.first[i] == size_t(-1) ? -1LL : (long long) p.first[i], (int) p.second[i]);
}

template <typename kmer_t, typename wrapper_t>
static int dump_path(wrapper_t wrapper, std::string path, int k, bool complements, bool lower_bound) {
    auto simplitigs = simplitigs_from_fasta(path);
    print_path(OverlapHamiltonianPath(wrapper, kmer_t(0), simplitigs, k, complements, lower_bound));
    return 0;
}

template <typename kmer_t, typename wrapper_t>
static int dump_sparse_path(wrapper_t wrapper, std::string path, int k, bool complements, bool lower_bound) {
    auto simplitigs = simplitigs_from_fasta(path);
    std::vector<kmer_t> kmers;
    for (auto &s : simplitigs) {
        if ((int) (s.size() / 2) != k) return 3;
        kmers.push_back(simplitig_first(kmer_t(0), s, k));
    }
    print_path(OverlapHamiltonianPathSparse(wrapper, kmers, k, complements, lower_bound));
    return 0;
}

template <typename kmer_t, typename wrapper_t>
static int run(wrapper_t wrapper, int argc, char **argv) {
    std::string cmd = argv[1], path = argv[2];
    int k = std::atoi(argv[3]);
    bool complements = std::atoi(argv[4]) != 0;
    if (cmd == "kmers" && argc == 6) return dump_kmers<kmer_t>(wrapper, path, k, complements, argv[5]);
    if (cmd == "full" && argc == 7) return run_full<kmer_t>(wrapper, path, k, complements, std::atoi(argv[5]), argv[6]);
    if (cmd == "path" && argc == 6) return dump_path<kmer_t>(wrapper, path, k, complements, std::atoi(argv[5]) != 0);
    if (cmd == "sparsepath" && argc == 6)
        return dump_sparse_path<kmer_t>(wrapper, path, k, complements, std::atoi(argv[5]) != 0);
    return 64;
}

int main(int argc, char **argv) {
    if (argc < 5) {
        std::fprintf(stderr, "usage: ref_harness kmers|path|sparsepath <fasta> <k> <complements> ...\n");
        return 64;
    }
    int k = std::atoi(argv[3]);
    if (k < 1 || k > 127) return 64;
    if (k < 32) return run<kmer64_t>(kmer_dict64_t(), argc, argv);
    if (k < 64) return run<kmer128_t>(kmer_dict128_t(), argc, argv);
    return run<kmer256_t>(kmer_dict256_t(), argc, argv);
}

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

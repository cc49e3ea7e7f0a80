// Host-side FASTA/FASTQ framing with the semantics of the reference's reader: kseq_read (src/kseq.h:182-224)
// driven by the loop of ReadKMers (src/parser.h:106-118), which stops at the first negative return value.
// Pure host C++ (no CUDA); part of libkcgpu so the CLI and the Python bindings frame files the same way.
//
// Output: the record sequences concatenated, each followed by one '\n' (a non-ACGT byte, so that k-mer windows
// never span records on the GPU), plus the offset and length of every record.
#include "../../include/kcgpu.h"

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <thread>
#include <vector>

namespace {

struct Reader {
    const uint8_t *d;
    uint64_t n, pos;
    int getc() { return pos < n ? d[pos++] : -1; }
    bool eof() const { return pos >= n; }
};

inline bool kseq_isspace(int c) { return c == ' ' || c == '\t' || c == '\n' || c == '\v' || c == '\f' || c == '\r'; }

// ks_getuntil2(KS_SEP_LINE, append = 1), src/kseq.h:98-147: append the rest of the line, drop one trailing '\r'
// when more than one character has been accumulated since `start`.  Returns false when already at EOF.
bool append_line(Reader &r, std::vector<uint8_t> &out, size_t start) {
    if (r.eof()) return false;
    const uint8_t *nl = (const uint8_t *) std::memchr(r.d + r.pos, '\n', r.n - r.pos);
    uint64_t end = nl ? (uint64_t) (nl - r.d) : r.n;
    out.insert(out.end(), r.d + r.pos, r.d + end);
    r.pos = nl ? end + 1 : r.n;
    if (out.size() - start > 1 && out.back() == '\r') out.pop_back();
    return true;
}

// Large plain FASTA files, framed by several threads.  When the file starts with '>', holds no '\r' and no line that starts with
// '+' or '@' (no FASTQ record, no stray marker), kseq_read reduces to: a line that starts with '>' is a header and ends the record
// before it; every other line is appended to the current record without its '\n' (empty lines add nothing).  Every line is owned by
// the byte range that holds its first byte, so the ranges are independent: pass 1 counts headers and sequence bytes per range and
// checks the three conditions, a prefix sum places every range's output, pass 2 copies the lines.  Anything else -> returns false and
// the serial reader below frames the file (3.1 GB took 3.2 s on one core: more than the GPUs need for the whole job).
struct FrameRange {
    uint64_t begin = 0, end = 0;     // first bytes of the lines this range owns lie in [begin, end)
    uint64_t headers = 0, bases = 0;  // lines starting with '>', bytes of the other lines (without '\n')
    bool plain = true;
};

inline uint64_t first_line_start(const uint8_t *d, uint64_t n, uint64_t from) {
    if (from == 0) return 0;
    const uint8_t *nl = (const uint8_t *) std::memchr(d + from - 1, '\n', n - (from - 1));
    return nl ? (uint64_t) (nl - d) + 1 : n;
}

bool frame_plain_fasta_parallel(const uint8_t *d, uint64_t n, uint8_t **seq_out, uint64_t *n_bytes, uint64_t **rec_off_out, uint64_t **rec_len_out,
                                uint64_t *n_recs, int *rc) {
    const uint64_t MIN_BYTES = 8ull << 20;
    unsigned hw = std::thread::hardware_concurrency();
    if (const char *e = std::getenv("KC_FRAME_THREADS")) hw = (unsigned) std::atoi(e);
    if (n < MIN_BYTES && !std::getenv("KC_FRAME_THREADS")) return false;
    if (hw < 2 || n < 2 || d[0] != '>') return false;
    const unsigned T = hw > 32 ? 32 : hw;
    std::vector<FrameRange> rg(T);
    for (unsigned t = 0; t < T; ++t) {
        rg[t].begin = first_line_start(d, n, n / T * t);
        rg[t].end = t + 1 < T ? first_line_start(d, n, n / T * (t + 1)) : n;
    }
    auto pass1 = [&](unsigned t) {
        FrameRange &g = rg[t];
        uint64_t p = g.begin;
        while (p < g.end) {
            const uint8_t *nl = (const uint8_t *) std::memchr(d + p, '\n', n - p);
            const uint64_t e = nl ? (uint64_t) (nl - d) : n;
            const uint8_t c = d[p];
            if (c == '>') {
                ++g.headers;
                if (!nl) g.plain = false;  // a header without its line end: kseq_read's corner cases
            } else if (c == '+' || c == '@') {
                g.plain = false;
            } else {
                g.bases += e - p;
            }
            if (std::memchr(d + p, '\r', e - p)) g.plain = false;
            p = e + 1;
        }
    };
    {
        std::vector<std::thread> th;
        for (unsigned t = 1; t < T; ++t) th.emplace_back(pass1, t);
        pass1(0);
        for (auto &x : th) x.join();
    }
    uint64_t recs = 0, bases = 0;
    for (unsigned t = 0; t < T; ++t) {
        if (!rg[t].plain) return false;
        recs += rg[t].headers;
        bases += rg[t].bases;
    }
    if (recs == 0) return false;
    const uint64_t total = bases + recs;  // one '\n' behind every record
    uint8_t *seq = (uint8_t *) std::malloc(total + 64);
    uint64_t *off = (uint64_t *) std::malloc(recs * 8 + 8), *len = (uint64_t *) std::malloc(recs * 8 + 8);
    if (!seq || !off || !len) {
        std::free(seq);
        std::free(off);
        std::free(len);
        *rc = KC_ERR_OOM;
        return true;
    }
    std::vector<uint64_t> rec0(T), base0(T);  // headers / bases before the range
    uint64_t rsum = 0, bsum = 0;
    for (unsigned t = 0; t < T; ++t) {
        rec0[t] = rsum;
        base0[t] = bsum;
        rsum += rg[t].headers;
        bsum += rg[t].bases;
    }
    auto pass2 = [&](unsigned t) {
        const FrameRange &g = rg[t];
        uint64_t r = rec0[t], b = base0[t];  // headers seen so far, bases copied so far: the next base goes to seq[b + r - 1]
        uint64_t p = g.begin;
        while (p < g.end) {
            const uint8_t *nl = (const uint8_t *) std::memchr(d + p, '\n', n - p);
            const uint64_t e = nl ? (uint64_t) (nl - d) : n;
            if (d[p] == '>') {
                if (r) seq[b + r - 1] = '\n';  // ends record r - 1
                off[r] = b + r;
                ++r;
            } else {
                std::memcpy(seq + b + r - 1, d + p, e - p);
                b += e - p;
            }
            p = e + 1;
        }
    };
    {
        std::vector<std::thread> th;
        for (unsigned t = 1; t < T; ++t) th.emplace_back(pass2, t);
        pass2(0);
        for (auto &x : th) x.join();
    }
    seq[total - 1] = '\n';
    for (uint64_t r = 0; r < recs; ++r) len[r] = (r + 1 < recs ? off[r + 1] : total) - off[r] - 1;
    *seq_out = seq;
    *n_bytes = total;
    *rec_off_out = off;
    *rec_len_out = len;
    *n_recs = recs;
    *rc = KC_OK;
    return true;
}

}  // namespace

extern "C" int kc_frame_fasta(const uint8_t *data, uint64_t n, uint8_t **seq_out, uint64_t *n_bytes, uint64_t **rec_off_out,
                              uint64_t **rec_len_out, uint64_t *n_recs) {
    if ((!data && n) || !seq_out || !n_bytes || !rec_off_out || !rec_len_out || !n_recs) return KC_ERR_ARG;
    try {
        int prc = KC_OK;
        if (frame_plain_fasta_parallel(data, n, seq_out, n_bytes, rec_off_out, rec_len_out, n_recs, &prc)) return prc;
        Reader r{data, n, 0};
        std::vector<uint8_t> seq;
        seq.reserve(n + 16);
        std::vector<uint64_t> off, len;
        std::vector<uint8_t> qual;
        int last_char = 0;
        while (true) {
            if (last_char == 0) {  // src/kseq.h:187-191: skip to the next '>' or '@', wherever it is
                int c;
                while ((c = r.getc()) >= 0 && c != '>' && c != '@') {}
                if (c < 0) break;
                last_char = c;
            }
            if (r.eof()) break;  // :193 the name read finds nothing -> -1
            int delim = 0;
            while (!r.eof()) {  // name: up to the first white-space character
                int c = r.getc();
                if (kseq_isspace(c)) {
                    delim = c;
                    break;
                }
            }
            if (delim != '\n') {  // :194 comment = rest of the header line
                const uint8_t *nl = r.eof() ? nullptr : (const uint8_t *) std::memchr(r.d + r.pos, '\n', r.n - r.pos);
                r.pos = nl ? (uint64_t) (nl - r.d) + 1 : r.n;
            }
            const size_t start = seq.size();
            int c;
            while ((c = r.getc()) >= 0 && c != '>' && c != '+' && c != '@') {  // :199-203
                if (c == '\n') continue;
                seq.push_back((uint8_t) c);
                append_line(r, seq, start);
            }
            if (c == '>' || c == '@') last_char = c;
            if (c != '+') {  // FASTA record
                off.push_back(start);
                len.push_back(seq.size() - start);
                seq.push_back('\n');
                continue;
            }
            while ((c = r.getc()) >= 0 && c != '\n') {}  // :217 rest of the '+' line
            const size_t seq_len = seq.size() - start;
            if (c < 0) {  // -2: no quality string; the record is dropped and reading stops
                seq.resize(start);
                break;
            }
            qual.clear();
            while (append_line(r, qual, 0) && qual.size() < seq_len) {}  // :219
            last_char = 0;
            if (qual.size() != seq_len) {  // -2: truncated quality
                seq.resize(start);
                break;
            }
            off.push_back(start);
            len.push_back(seq_len);
            seq.push_back('\n');
        }
        if (seq.empty()) seq.push_back('\n');
        *n_bytes = seq.size();
        *n_recs = off.size();
        *seq_out = (uint8_t *) std::malloc(seq.size() + 64);
        *rec_off_out = (uint64_t *) std::malloc(off.size() * 8 + 8);
        *rec_len_out = (uint64_t *) std::malloc(len.size() * 8 + 8);
        if (!*seq_out || !*rec_off_out || !*rec_len_out) return KC_ERR_OOM;
        std::memcpy(*seq_out, seq.data(), seq.size());
        if (!off.empty()) {
            std::memcpy(*rec_off_out, off.data(), off.size() * 8);
            std::memcpy(*rec_len_out, len.data(), len.size() * 8);
        }
        return KC_OK;
    } catch (const std::bad_alloc &) {
        return KC_ERR_OOM;
    }
}

// ---- masked-superstring text conversions (reference src/conversions.h; host only) ----------------------
namespace {
inline bool ms_is_upper(uint8_t c) { return c <= 'Z'; }                                         // src/conversions.h:6-8
inline uint8_t ms_to_upper(uint8_t c) { return ms_is_upper(c) ? c : (uint8_t) (c - ('a' - 'A')); }  // :10-14
inline uint8_t ms_masked(uint8_t c, bool mask) {                                                // src/kmers.h:124-127
    const int d = (int) (c <= 'Z') - (int) mask;
    return (uint8_t) (c + d * ('a' - 'A'));
}
int give(const std::vector<uint8_t> &v, uint8_t **out, uint64_t *n_out) {
    *out = (uint8_t *) std::malloc(v.size() + 1);
    if (!*out) return KC_ERR_OOM;
    if (!v.empty()) std::memcpy(*out, v.data(), v.size());
    (*out)[v.size()] = 0;
    *n_out = v.size();
    return KC_OK;
}
}  // namespace

// split_ms, src/conversions.h:16-33: mask as '0'/'1' characters, superstring in upper case (no newlines added here).
extern "C" int kc_split_ms(const uint8_t *ms, uint64_t n, uint8_t **superstring, uint8_t **mask) {
    if ((!ms && n) || !superstring || !mask) return KC_ERR_ARG;
    *superstring = (uint8_t *) std::malloc(n + 1);
    *mask = (uint8_t *) std::malloc(n + 1);
    if (!*superstring || !*mask) return KC_ERR_OOM;
    for (uint64_t i = 0; i < n; ++i) {
        (*mask)[i] = ms_is_upper(ms[i]) ? '1' : '0';
        (*superstring)[i] = ms_to_upper(ms[i]);
    }
    (*superstring)[n] = (*mask)[n] = 0;
    return KC_OK;
}

// join_ms, src/conversions.h:35-44 (the text after ">superstring\n"): a missing mask character counts as '0'.
extern "C" int kc_join_ms(const uint8_t *superstring, uint64_t n_s, const uint8_t *mask, uint64_t n_m, uint8_t **out, uint64_t *n_out) {
    if ((!superstring && n_s) || (!mask && n_m) || !out || !n_out) return KC_ERR_ARG;
    try {
        std::vector<uint8_t> v(n_s);
        for (uint64_t i = 0; i < n_s; ++i) v[i] = ms_masked(superstring[i], i < n_m && mask[i] == '1');
        return give(v, out, n_out);
    } catch (const std::bad_alloc &) {
        return KC_ERR_OOM;
    }
}

// ms_to_spss, src/conversions.h:46-72: the complete FASTA text.  Each maximal run of upper-case letters plus the k-1
// letters after it (upper-cased, clipped at the end) is one record ">i"; a run that reaches the end of the superstring
// is left without its newline, exactly as the reference prints it.
extern "C" int kc_ms_to_spss(const uint8_t *ms, uint64_t n, int k, uint8_t **out, uint64_t *n_out) {
    if ((!ms && n) || !out || !n_out || k < 1) return KC_ERR_ARG;
    try {
        std::vector<uint8_t> v;
        v.reserve(n + n / 8 + 64);
        bool masked = false;
        uint64_t counter = 0;
        char num[32];
        for (uint64_t i = 0; i < n; ++i) {
            if (ms_is_upper(ms[i])) {
                if (!masked) {
                    const int w = std::snprintf(num, sizeof(num), ">%llu\n", (unsigned long long) counter++);
                    v.insert(v.end(), num, num + w);
                }
                masked = true;
                v.push_back(ms[i]);
            } else {
                if (!masked) continue;
                masked = false;
                for (int j = 0; j < k - 1; ++j)
                    if (i + (uint64_t) j < n) v.push_back(ms_to_upper(ms[i + j]));
                v.push_back('\n');
            }
        }
        return give(v, out, n_out);
    } catch (const std::bad_alloc &) {
        return KC_ERR_OOM;
    }
}

// spss_to_ms, src/conversions.h:74-91 (the text after the header line): records shorter than k are skipped, the last
// k-1 letters of every record are OFF.
extern "C" int kc_spss_to_ms(const uint8_t *seq, const uint64_t *rec_off, const uint64_t *rec_len, uint64_t n_recs, int k, uint8_t **out,
                             uint64_t *n_out) {
    if (!out || !n_out || k < 1 || (n_recs && (!seq || !rec_off || !rec_len))) return KC_ERR_ARG;
    try {
        std::vector<uint8_t> v;
        for (uint64_t r = 0; r < n_recs; ++r) {
            const uint64_t l = rec_len[r];
            if (l < (uint64_t) k) continue;
            const uint8_t *s = seq + rec_off[r];
            for (uint64_t i = 0; i < l; ++i) v.push_back(ms_masked(s[i], i + (uint64_t) k <= l));
        }
        return give(v, out, n_out);
    } catch (const std::bad_alloc &) {
        return KC_ERR_OOM;
    }
}

// Name and comment of the FIRST record as kseq_read parses them (src/kseq.h:187-194), for ReprintSequenceHeader
// (src/masks.h:27-37).  span = {name_off, name_len, has_comment, comment_off, comment_len}; all zero without a record.
extern "C" int kc_fasta_first_header(const uint8_t *data, uint64_t n, uint64_t *span) {
    if ((!data && n) || !span) return KC_ERR_ARG;
    for (int i = 0; i < 5; ++i) span[i] = 0;
    uint64_t p = 0;
    while (p < n && data[p] != '>' && data[p] != '@') ++p;
    if (p >= n) return KC_OK;
    ++p;
    span[0] = p;
    while (p < n && !kseq_isspace(data[p])) ++p;
    span[1] = p - span[0];
    if (p >= n || data[p] == '\n') return KC_OK;
    ++p;  // the delimiter
    span[2] = 1;
    span[3] = p;
    const uint8_t *nl = p < n ? (const uint8_t *) std::memchr(data + p, '\n', n - p) : nullptr;
    uint64_t end = nl ? (uint64_t) (nl - data) : n;
    if (end - p > 1 && data[end - 1] == '\r') --end;
    span[4] = end - p;
    return KC_OK;
}

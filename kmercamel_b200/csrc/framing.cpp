// Host-side FASTA/FASTQ framing with the semantics of the reference's reader: kseq_read (src/kseq.h:182-224)
// driven by the loop of ReadKMers (src/parser.h:106-118), which stops at the first negative return value.
// Pure host C++ (no CUDA); part of libkcgpu so the CLI and the Python bindings frame files the same way.
//
// Output: the record sequences concatenated, each followed by one '\n' (a non-ACGT byte, so that k-mer windows
// never span records on the GPU), plus the offset and length of every record.
#include "../../include/kcgpu.h"

#include <cstdlib>
#include <cstring>
#include <new>
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

}  // namespace

extern "C" int kc_frame_fasta(const uint8_t *data, uint64_t n, uint8_t **seq_out, uint64_t *n_bytes, uint64_t **rec_off_out,
                              uint64_t **rec_len_out, uint64_t *n_recs) {
    if ((!data && n) || !seq_out || !n_bytes || !rec_off_out || !rec_len_out || !n_recs) return KC_ERR_ARG;
    try {
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

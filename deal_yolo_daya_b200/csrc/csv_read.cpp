// Native CSV ingest: the tokenizer behind `pd.read_csv(path, encoding="utf-8[-sig]")`
// (processor.py:124-128, 181-182, 235, 379) for the text columns of the pipeline, multi-threaded.
//
// It restates the state machine of pandas' C tokenizer (pandas/_libs/src/parser/tokenizer.c,
// tokenize_bytes) for the options the reference uses -- delimiter ',', quotechar '"', doublequote,
// no escapechar / comment, skip_blank_lines, \n and \r\n terminators:
//   * a quote opens a quoted field only as the first byte of a field; inside an unquoted field it is
//     an ordinary byte; after the closing quote of a quoted field further bytes are appended ("ab"c -> abc)
//   * "" inside a quoted field is one quote; terminators inside a quoted field are data
//   * empty and whitespace-only (space / tab) lines are skipped; a UTF-8 BOM at offset 0 is skipped
//   * cells equal to one of pandas' NA strings (passed in by the caller) are missing values
// What it does NOT decide is handed back to pandas: header names (the caller parses the header record
// with pandas), dtype inference of columns that are not certainly text (the caller re-parses those
// columns with pandas from the cell texts this file extracts), and every input with a feature
// outside the list above (bare \r terminators, NUL bytes, ragged rows, EOF inside a quoted field,
// invalid UTF-8): those set FLAG_UNSUPPORTED and the caller runs pd.read_csv on the whole file.
//
// Parallel tokenisation: the input is cut at line feeds into one chunk per thread; a chunk is
// scanned under both hypotheses for its first byte (outside / inside a quoted field) and the chunks
// are stitched left to right, which fixes the true hypothesis of each.
#include <fcntl.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <chrono>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/dyd.h"
#include "simd_text.hpp"

namespace {

enum : int32_t { FLAG_UNSUPPORTED = 1, FLAG_EMPTY = 2, FLAG_WIDE = 4 };

template <typename F>
void parallel_ranges(int64_t n, int n_threads, int64_t grain, F f) {
    if (n_threads <= 0) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
    n_threads = (int)std::min<int64_t>(n_threads, std::max<int64_t>(1, n / grain));
    if (n_threads <= 1) { f(0, n, 0); return; }
    std::vector<std::thread> th;
    const int64_t per = (n + n_threads - 1) / n_threads;
    for (int t = 0; t < n_threads; ++t) {
        const int64_t a = t * per, b = std::min(n, a + per);
        if (a < b) th.emplace_back([=] { f(a, b, t); });
    }
    for (auto& t : th) t.join();
}

// ---- phase A: line feeds that terminate a record -------------------------------------------------
enum QState : uint8_t { FIELD_START, IN_FIELD, IN_QUOTED, QUOTE_IN_QUOTED };

struct ChunkScan {
    std::vector<int64_t> lf;     // positions of record-terminating '\n' inside the chunk
    bool ends_in_quote = false;  // state after the chunk's last byte is inside a quoted field
    bool bad = false;            // bare '\r' outside quotes or NUL byte
};

void scan_chunk(const uint8_t* d, int64_t begin, int64_t end, int64_t n, bool start_in_quote, ChunkScan& out) {
    QState st = start_in_quote ? IN_QUOTED : FIELD_START;
    int64_t i = begin;
    while (i < end) {
        if (st == IN_QUOTED) {
            const void* q = memchr(d + i, '"', (size_t)(end - i));
            if (memchr(d + i, 0, (size_t)((q ? (const uint8_t*)q : d + end) - (d + i)))) out.bad = true;
            if (!q) { i = end; break; }
            i = (const uint8_t*)q - d + 1;
            st = QUOTE_IN_QUOTED;
            continue;
        }
        const uint8_t c = d[i];
        if (st == QUOTE_IN_QUOTED && c == '"') { st = IN_QUOTED; ++i; continue; }
        switch (c) {
            case ',': st = FIELD_START; break;
            case '\n': out.lf.push_back(i); st = FIELD_START; break;
            case '\r':
                if (i + 1 >= n || d[i + 1] != '\n') out.bad = true;       // bare CR terminator: not restated here
                st = IN_FIELD;                                              // the following '\n' ends the record
                break;
            case '"': st = (st == FIELD_START) ? IN_QUOTED : IN_FIELD; break;
            case 0: out.bad = true; st = IN_FIELD; break;
            default: st = IN_FIELD; break;
        }
        ++i;
    }
    out.ends_in_quote = (st == IN_QUOTED);
}

// ---- one record ---------------------------------------------------------------------------------
// Walks the record that starts at `p` (state START_FIELD) and ends at `e` (exclusive: the position of
// its terminating '\n', or the end of the data).  A '\r' directly before `e` belongs to the terminator.
// F(field_index, ptr, len) receives raw pieces of a field; a field may arrive in several pieces
// (around "" and around the closing quote).  G(field_index) closes a field.  Returns the field count,
// or -1 when the data ends inside a quoted field.
template <typename Piece, typename Close>
inline int walk_record(const uint8_t* p, const uint8_t* e, Piece piece, Close close) {
    if (e > p && e[-1] == '\r') --e;
    int f = 0;
    for (;;) {
        // START_FIELD
        if (p < e && *p == '"') {
            ++p;
            for (;;) {                                   // IN_QUOTED_FIELD
                const uint8_t* q = (const uint8_t*)memchr(p, '"', (size_t)(e - p));
                if (!q) return -1;
                if (q > p) piece(f, p, (int64_t)(q - p));
                p = q + 1;
                if (p < e && *p == '"') { piece(f, p, 1); ++p; continue; }      // "" -> one quote
                break;
            }
            // QUOTE_IN_QUOTED_FIELD followed by anything but a delimiter: the rest is an unquoted tail
        }
        const uint8_t* q = (const uint8_t*)memchr(p, ',', (size_t)(e - p));
        const uint8_t* stop = q ? q : e;
        if (stop > p) piece(f, p, (int64_t)(stop - p));
        close(f);
        ++f;
        if (!q) return f;
        p = q + 1;
    }
}

inline bool is_blank(const uint8_t* p, const uint8_t* e) {
    if (e > p && e[-1] == '\r') --e;
    for (; p < e; ++p) if (*p != ' ' && *p != '\t') return false;
    return true;
}

// strict UTF-8 (what CPython's decoder and Arrow accept): no overlongs, no surrogates, <= U+10FFFF
inline bool utf8_ok(const uint8_t* s, int64_t n) {
    int64_t i = 0;
    while (i < n) {
        if (i + 8 <= n) {                                 // ASCII fast path
            uint64_t w;
            memcpy(&w, s + i, 8);
            if (!(w & 0x8080808080808080ULL)) { i += 8; continue; }
        }
        const uint8_t c = s[i];
        if (c < 0x80) { ++i; continue; }
        int len;
        uint32_t cp;
        if ((c & 0xE0) == 0xC0) { len = 2; cp = c & 0x1F; }
        else if ((c & 0xF0) == 0xE0) { len = 3; cp = c & 0x0F; }
        else if ((c & 0xF8) == 0xF0) { len = 4; cp = c & 0x07; }
        else return false;
        if (i + len > n) return false;
        for (int k = 1; k < len; ++k) {
            if ((s[i + k] & 0xC0) != 0x80) return false;
            cp = (cp << 6) | (s[i + k] & 0x3F);
        }
        if ((len == 2 && cp < 0x80) || (len == 3 && cp < 0x800) || (len == 4 && cp < 0x10000)) return false;
        if (cp > 0x10FFFF || (cp >= 0xD800 && cp <= 0xDFFF)) return false;
        i += len;
    }
    return true;
}

// A cell pandas cannot turn into a number, a boolean or a missing value: its first byte cannot start
// any of those literals.  (Conservative: anything else is "uncertain" and goes back to pandas.)
inline bool certainly_text(const uint8_t* s, int64_t n) {
    if (n == 0) return false;
    const uint8_t c = s[0];
    if (c >= 0x80 || c == '{' || c == '[') return true;
    if ((c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z')) {
        switch (c | 0x20) { case 'i': case 'n': case 't': case 'f': case 'e': case 'j': return false; default: return true; }
    }
    return false;
}

struct Csv {
    const uint8_t* d = nullptr;
    int64_t n = 0;
    std::vector<std::string> na;                 // pandas' NA strings
    int64_t header_begin = 0, header_end = 0;    // header record, terminator excluded
    std::vector<int64_t> row_begin, row_end;     // data records (non-blank lines after the header)
    int32_t n_cols = 0, flags = 0;
    // measure()
    std::vector<std::vector<int64_t>> off;       // per column: n_rows + 1 offsets of the unescaped cell text
    std::vector<std::vector<uint8_t>> isna;      // per column: 1 = missing
    // wide (AVX-512) tokenizer: end position and quote count of every field, row-major [n_rows * n_cols]
    bool fast = false;
    std::vector<int64_t> fend;
    std::vector<uint32_t> fquotes;

    bool is_na(const uint8_t* s, int64_t len) const {
        if (len == 0) return true;
        if (len > 9) return false;
        for (const auto& v : na) if ((int64_t)v.size() == len && memcmp(v.data(), s, (size_t)len) == 0) return true;
        return false;
    }
};


// ---- wide tokenizer ------------------------------------------------------------------------------
// Same records and fields as the state machine above for the files it accepts, found with 64-byte compares:
// quote parity by carry-less multiplication (a byte is inside a quoted field iff an odd number of quotes precede or
// sit on it), separators = commas / line feeds outside quotes.  That shortcut is only the pandas tokenizer when
// quotes appear nowhere but around whole fields and doubled inside them, so the scan CHECKS that: every opening
// quote follows a separator, the start of the data or a closing quote (""), every closing quote is followed by a
// separator, a quote or the end of the data; no '\r', no NUL.  Any other file makes the scan give up and
// dyd_csv_open falls back to the byte-wise scanner (which knows `ab"c` and `"ab"c`).
#define DYD_TOK __attribute__((target("avx512f,avx512bw,avx512vl,avx512vbmi2,bmi2,popcnt,pclmul")))

struct FastChunk {
    int64_t quotes = 0;          // pass 1
    bool bad = false;
    std::vector<int64_t> sep;    // pass 2: separator positions, line feeds flagged in bit 62
    std::vector<uint32_t> sepq;  // quotes between the previous separator (or the chunk start) and this one
    uint32_t tail_quotes = 0;    // quotes after the last separator of the chunk
    bool first_open_needs_prev = false;   // the chunk's first byte is an opening quote: the previous byte decides
    bool last_close_needs_next = false;   // the chunk's last byte is a closing quote: the next byte decides
};
constexpr int64_t LF_FLAG = 1LL << 62;

DYD_TOK inline uint64_t prefix_xor(uint64_t q) {
    const __m128i r = _mm_clmulepi64_si128(_mm_set_epi64x(0, (long long)q), _mm_set1_epi8((char)0xFF), 0);
    return (uint64_t)_mm_cvtsi128_si64(r);
}

DYD_TOK void fast_pass1(const uint8_t* d, int64_t a, int64_t b, FastChunk& out) {
    const __m512i vq = _mm512_set1_epi8('"'), vr = _mm512_set1_epi8('\r'), vz = _mm512_setzero_si512();
    int64_t q = 0; uint64_t bad = 0;
    int64_t i = a;
    for (; i + 64 <= b; i += 64) {
        const __m512i v = _mm512_loadu_si512(d + i);
        q += _mm_popcnt_u64(_mm512_cmpeq_epi8_mask(v, vq));
        bad |= _mm512_cmpeq_epi8_mask(v, vr) | _mm512_cmpeq_epi8_mask(v, vz);
    }
    if (i < b) {
        const __mmask64 m = (~0ULL) >> (64 - (b - i));
        const __m512i v = _mm512_maskz_loadu_epi8(m, d + i);
        q += _mm_popcnt_u64(_mm512_cmpeq_epi8_mask(v, vq));
        bad |= (_mm512_cmpeq_epi8_mask(v, vr) | _mm512_cmpeq_epi8_mask(v, vz)) & m;
    }
    out.quotes = q; out.bad = bad != 0;
}

// [a, b): a is a multiple of 64 relative to `base` is NOT required; `inside` = state before byte a
DYD_TOK void fast_pass2(const uint8_t* d, int64_t a, int64_t b, int64_t n, int64_t data_start, bool inside, FastChunk& out) {
    const __m512i vq = _mm512_set1_epi8('"'), vc = _mm512_set1_epi8(','), vn = _mm512_set1_epi8('\n');
    uint64_t carry = inside ? ~0ULL : 0ULL;       // parity before the block, spread over all bits
    uint64_t prev_allow = 0;                      // bit 0: the byte before this block allows an opening quote
    bool have_prev = false;                       // false for the first block of the chunk (decided by the caller)
    uint64_t pend_close = 0;                      // the previous block's last byte was a closing quote
    uint32_t run_q = 0;                           // quotes since the last separator
    out.sep.reserve((size_t)((b - a) / 2048 + 16));
    out.sepq.reserve((size_t)((b - a) / 2048 + 16));
    for (int64_t i = a; i < b; i += 64) {
        const int64_t len = std::min<int64_t>(64, b - i);
        const __mmask64 m = len == 64 ? ~0ULL : ((~0ULL) >> (64 - len));
        const __m512i v = _mm512_maskz_loadu_epi8(m, d + i);
        const uint64_t q = _mm512_cmpeq_epi8_mask(v, vq) & m;
        const uint64_t cm = _mm512_cmpeq_epi8_mask(v, vc) & m, lf = _mm512_cmpeq_epi8_mask(v, vn) & m;
        const uint64_t P = prefix_xor(q) ^ carry;               // bit k: inside a quoted field after byte k
        const uint64_t S = (cm | lf) & ~P;                       // separators
        const uint64_t O = q & P, C = q & ~P;                   // opening / closing quotes
        // a closing quote needs a separator, a quote or the end of the data behind it
        const uint64_t follow = cm | lf | q;
        if (pend_close && !(follow & 1)) { out.bad = true; return; }
        uint64_t need_next = C & ~(follow >> 1);
        pend_close = 0;
        if (need_next & (1ULL << (len - 1))) { need_next &= ~(1ULL << (len - 1)); pend_close = 1; }
        if (need_next) { out.bad = true; return; }
        // an opening quote needs a separator, the start of the data or a closing quote in front of it
        const uint64_t allow = ((S | C) << 1) | (have_prev ? prev_allow : 0);
        uint64_t bad_open = O & ~allow;
        if (!have_prev && (bad_open & 1)) {                    // first byte of the chunk: the caller checks the byte before it
            bad_open &= ~1ULL;
            if (i == data_start) { /* start of the data: allowed */ } else out.first_open_needs_prev = true;
        }
        if (bad_open) { out.bad = true; return; }
        prev_allow = ((S | C) >> (len - 1)) & 1;
        have_prev = true;
        // separators out, with the quotes seen since the previous one
        uint64_t s = S;
        uint64_t done = 0;                                       // bits below the current separator
        while (s) {
            const int k = __builtin_ctzll(s);
            const uint64_t below = (k == 63) ? ~0ULL : ((1ULL << (k + 1)) - 1);
            run_q += (uint32_t)_mm_popcnt_u64(q & below & ~done);
            out.sep.push_back((i + k) | (((lf >> k) & 1) ? LF_FLAG : 0));
            out.sepq.push_back(run_q);
            run_q = 0;
            done = below;
            s &= s - 1;
        }
        run_q += (uint32_t)_mm_popcnt_u64(q & ~done);
        carry = (uint64_t)0 - ((P >> (len - 1)) & 1);
    }
    out.tail_quotes = run_q;
    out.last_close_needs_next = pend_close != 0;
    (void)n;
}

// unescaped copy of the inside of a quoted field: every second quote of a run of quotes is dropped
DYD_TOK inline uint8_t* unescape_wide(uint8_t* o, const uint8_t* s, int64_t n) {
    const __m512i vq = _mm512_set1_epi8('"');
    const uint64_t EVEN = 0x5555555555555555ULL;
    uint64_t pending = 0;                                        // the previous block ended on the first quote of a pair
    for (int64_t i = 0; i < n; i += 64) {
        const int64_t len = std::min<int64_t>(64, n - i);
        const __mmask64 m = len == 64 ? ~0ULL : ((~0ULL) >> (64 - len));
        const __m512i v = _mm512_maskz_loadu_epi8(m, s + i);
        uint64_t q = _mm512_cmpeq_epi8_mask(v, vq) & m;
        uint64_t drop = 0;
        if (pending) { drop = q & 1; q &= ~1ULL; }
        if (q) {
            const uint64_t starts = q & ~(q << 1);
            const uint64_t se = starts & EVEN;
            const uint64_t re = q & ~(q + se);                   // runs that start on an even position
            const uint64_t ro = q & ~re;
            drop |= (re & ~EVEN) | (ro & EVEN);
        }
        const uint64_t last = 1ULL << (len - 1);
        pending = ((q | (drop & 1)) & last) && !(drop & last);
        const uint64_t keep = m & ~drop;
        const int cnt = (int)_mm_popcnt_u64(keep);
        _mm512_mask_storeu_epi8(o, cnt == 64 ? ~0ULL : ((1ULL << cnt) - 1), _mm512_maskz_compress_epi8(keep, v));
        o += cnt;
    }
    return o;
}

inline uint8_t* unescape_scalar(uint8_t* o, const uint8_t* s, int64_t n) {
    for (int64_t i = 0; i < n; ++i) { *o++ = s[i]; if (s[i] == '"') ++i; }
    return o;
}

DYD_TOK inline bool any_high_bit(const uint8_t* s, int64_t n) {
    int64_t i = 0;
    uint64_t acc = 0;
    for (; i + 64 <= n; i += 64) acc |= _mm512_movepi8_mask(_mm512_loadu_si512(s + i));
    if (i < n) acc |= _mm512_movepi8_mask(_mm512_maskz_loadu_epi8((~0ULL) >> (64 - (n - i)), s + i));
    return acc != 0;
}

inline bool fast_any_high_bit(const uint8_t* s, int64_t n) { return any_high_bit(s, n); }
inline void fast_unescape(uint8_t* o, const uint8_t* s, int64_t n) { unescape_wide(o, s, n); }

// Fills header / rows / field table of `c`; false = not a file for this scanner (nothing in `c` is touched then).
static double tnow() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
bool fast_tokenize(Csv* c, int64_t start, int threads) {
    const bool timing = getenv("DYD_CSV_TIMING") != nullptr;
    double t0 = tnow();
    auto lap = [&](const char* what) { if (timing) { double t1 = tnow(); fprintf(stderr, "  csv wide %-10s %.4f s\n", what, t1 - t0); t0 = t1; } };
    if (!dyd_simd::wide() || !__builtin_cpu_supports("pclmul")) return false;
    if (const char* e = getenv("DYD_CSV_WIDE")) if (*e == '0') return false;       // tests compare the two scanners
    const uint8_t* d = c->d;
    const int64_t n = c->n;
    if (n - start < 1) return false;
    int64_t chunk_min = 1 << 20;
    if (const char* e = getenv("DYD_CSV_CHUNK_MIN")) chunk_min = std::max<int64_t>(1, atoll(e));   // tests: many chunks on small inputs
    int T = threads > 0 ? threads : (int)std::max(1u, std::thread::hardware_concurrency());
    T = (int)std::min<int64_t>(T, std::max<int64_t>(1, (n - start) / chunk_min));
    std::vector<int64_t> cut((size_t)T + 1);
    for (int t = 0; t <= T; ++t) cut[(size_t)t] = t == T ? n : start + (n - start) / T * t;
    std::vector<FastChunk> ch((size_t)T);
    {
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t) th.emplace_back([&, t] { fast_pass1(d, cut[(size_t)t], cut[(size_t)t + 1], ch[(size_t)t]); });
        for (auto& x : th) x.join();
    }
    lap("pass1");
    std::vector<uint8_t> inside((size_t)T + 1, 0);
    for (int t = 0; t < T; ++t) { if (ch[(size_t)t].bad) return false; inside[(size_t)t + 1] = inside[(size_t)t] ^ (uint8_t)(ch[(size_t)t].quotes & 1); }
    if (inside[(size_t)T]) return false;                         // the data ends inside a quoted field
    {
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t)
            th.emplace_back([&, t] { fast_pass2(d, cut[(size_t)t], cut[(size_t)t + 1], n, start, inside[(size_t)t] != 0, ch[(size_t)t]); });
        for (auto& x : th) x.join();
    }
    for (int t = 0; t < T; ++t) {
        const FastChunk& k = ch[(size_t)t];
        if (k.bad) return false;
        const int64_t a = cut[(size_t)t], b = cut[(size_t)t + 1];
        if (k.first_open_needs_prev) {                           // d[a] opens a field: d[a-1] must be a separator outside quotes or a closing quote
            const uint8_t pb = d[a - 1];
            const bool sep_before = (pb == ',' || pb == '\n') && !inside[(size_t)t];
            const bool close_before = pb == '"' && !inside[(size_t)t];     // state before a is "outside": that quote closed a field
            if (!sep_before && !close_before) return false;
        }
        if (k.last_close_needs_next && b < n) { const uint8_t nb = d[b]; if (nb != ',' && nb != '\n' && nb != '"') return false; }
    }
    lap("pass2");
    // lines -> records
    struct Line { int64_t begin, end; int64_t first_sep; int32_t n_sep; };          // n_sep = commas in the line
    std::vector<int64_t> sep; std::vector<uint32_t> sepq;
    {
        size_t total = 0;
        for (auto& k : ch) total += k.sep.size();
        sep.reserve(total + 1); sepq.reserve(total + 1);
        uint32_t carry_q = 0;
        for (auto& k : ch) {
            for (size_t i = 0; i < k.sep.size(); ++i) { sep.push_back(k.sep[i]); sepq.push_back(k.sepq[i] + (i == 0 ? carry_q : 0)); }
            carry_q = k.sep.empty() ? carry_q + k.tail_quotes : k.tail_quotes;
            std::vector<int64_t>().swap(k.sep); std::vector<uint32_t>().swap(k.sepq);
        }
        int64_t after_last_lf = start;                           // first byte behind the last record-terminating line feed
        for (size_t i = sep.size(); i-- > 0;) if (sep[i] & LF_FLAG) { after_last_lf = (sep[i] & ~LF_FLAG) + 1; break; }
        if (after_last_lf < n) { sep.push_back(n | LF_FLAG); sepq.push_back(carry_q); }   // the end of the data ends the last record
    }
    lap("merge");
    std::vector<Line> lines;
    {
        int64_t begin = start; int64_t first = 0; int32_t commas = 0;
        for (size_t i = 0; i < sep.size(); ++i) {
            if (sep[i] & LF_FLAG) {
                const int64_t e = sep[i] & ~LF_FLAG;
                lines.push_back(Line{begin, e, first, commas});
                begin = e + 1; first = (int64_t)i + 1; commas = 0;
            } else ++commas;
        }
    }
    int64_t hdr = -1;
    std::vector<uint8_t> keep(lines.size());
    for (size_t k = 0; k < lines.size(); ++k) {
        keep[k] = lines[k].end > lines[k].begin && !is_blank(d + lines[k].begin, d + lines[k].end);
        if (keep[k] && hdr < 0) hdr = (int64_t)k;
    }
    if (hdr < 0) return false;
    const int32_t nc = lines[(size_t)hdr].n_sep + 1;
    size_t nr = 0;
    for (size_t k = (size_t)hdr + 1; k < lines.size(); ++k) if (keep[k]) { if (lines[k].n_sep + 1 != nc) return false; ++nr; }
    if (nr == 0) return false;
    c->header_begin = lines[(size_t)hdr].begin; c->header_end = lines[(size_t)hdr].end;
    c->n_cols = nc;
    c->row_begin.resize(nr); c->row_end.resize(nr);
    c->fend.resize(nr * (size_t)nc); c->fquotes.resize(nr * (size_t)nc);
    size_t r = 0;
    for (size_t k = (size_t)hdr + 1; k < lines.size(); ++k) {
        if (!keep[k]) continue;
        c->row_begin[r] = lines[k].begin; c->row_end[r] = lines[k].end;
        for (int32_t j = 0; j < nc; ++j) {
            c->fend[r * (size_t)nc + (size_t)j] = sep[(size_t)lines[k].first_sep + (size_t)j] & ~LF_FLAG;
            c->fquotes[r * (size_t)nc + (size_t)j] = sepq[(size_t)lines[k].first_sep + (size_t)j];
        }
        ++r;
    }
    lap("rows");
    c->fast = true;
    c->flags |= FLAG_WIDE;
    return true;
}

}  // namespace

extern "C" int dyd_csv_open(const uint8_t* data, int64_t n, const uint8_t* na_bytes, const int64_t* na_off, int32_t n_na,
                            int32_t threads, void** handle) {
    if (!handle || n < 0 || (n > 0 && !data) || n_na < 0 || (n_na > 0 && (!na_bytes || !na_off))) return DYD_E_ARG;
    Csv* c = new Csv();
    c->d = data; c->n = n;
    for (int32_t k = 0; k < n_na; ++k) c->na.emplace_back((const char*)na_bytes + na_off[k], (size_t)(na_off[k + 1] - na_off[k]));
    *handle = c;
    const uint8_t* d = data;
    int64_t start = (n >= 3 && d[0] == 0xEF && d[1] == 0xBB && d[2] == 0xBF) ? 3 : 0;

    if (fast_tokenize(c, start, threads)) return 0;

    // ---- phase A: chunk boundaries right after a '\n', two hypotheses per chunk, stitch
    int T = threads > 0 ? threads : (int)std::max(1u, std::thread::hardware_concurrency());
    T = (int)std::min<int64_t>(T, std::max<int64_t>(1, (n - start) / (1 << 20)));
    std::vector<int64_t> cut(T + 1, n);
    cut[0] = start;
    for (int t = 1; t < T; ++t) {
        int64_t pos = start + (n - start) / T * t;
        pos = std::max(pos, cut[t - 1]);
        const void* q = pos < n ? memchr(d + pos, '\n', (size_t)(n - pos)) : nullptr;
        cut[t] = q ? (const uint8_t*)q - d + 1 : n;
    }
    std::vector<ChunkScan> out_u(T), out_q(T);
    {
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t)
            th.emplace_back([&, t] {
                if (cut[t] >= cut[t + 1]) return;
                scan_chunk(d, cut[t], cut[t + 1], n, false, out_u[t]);
                if (t > 0) scan_chunk(d, cut[t], cut[t + 1], n, true, out_q[t]);
            });
        for (auto& t : th) t.join();
    }
    std::vector<int64_t> lf;
    bool in_quote = false, bad = false;
    for (int t = 0; t < T; ++t) {
        if (cut[t] >= cut[t + 1]) continue;
        const ChunkScan& s = in_quote ? out_q[t] : out_u[t];
        lf.insert(lf.end(), s.lf.begin(), s.lf.end());
        bad |= s.bad;
        in_quote = s.ends_in_quote;
    }
    if (in_quote) bad = true;                                     // EOF inside a quoted field: pandas raises
    if (bad) { c->flags |= FLAG_UNSUPPORTED; return 0; }

    // ---- lines -> records: drop blank lines, first record is the header, count fields
    const int64_t n_lines = (int64_t)lf.size() + ((lf.empty() ? start : lf.back() + 1) < n ? 1 : 0);
    auto line_begin = [&](int64_t k) { return k == 0 ? start : lf[k - 1] + 1; };
    auto line_end = [&](int64_t k) { return k < (int64_t)lf.size() ? lf[k] : n; };
    std::vector<uint8_t> keep((size_t)n_lines);
    std::vector<int32_t> nf((size_t)n_lines);
    parallel_ranges(n_lines, threads, 4096, [&](int64_t a, int64_t b, int) {
        for (int64_t k = a; k < b; ++k) {
            const uint8_t *p = d + line_begin(k), *e = d + line_end(k);
            keep[k] = !is_blank(p, e);
            nf[k] = keep[k] ? walk_record(p, e, [](int, const uint8_t*, int64_t) {}, [](int) {}) : 0;
        }
    });
    int64_t hdr = -1;
    for (int64_t k = 0; k < n_lines; ++k) if (keep[k]) { hdr = k; break; }
    if (hdr < 0) { c->flags |= FLAG_EMPTY | FLAG_UNSUPPORTED; return 0; }
    c->header_begin = line_begin(hdr);
    c->header_end = line_end(hdr);
    if (c->header_end > c->header_begin && d[c->header_end - 1] == '\r') --c->header_end;
    c->n_cols = nf[hdr];
    c->row_begin.reserve((size_t)(n_lines - hdr));
    c->row_end.reserve((size_t)(n_lines - hdr));
    for (int64_t k = hdr + 1; k < n_lines; ++k) {
        if (!keep[k]) continue;
        if (nf[k] != c->n_cols) { c->flags |= FLAG_UNSUPPORTED; return 0; }     // ragged (or EOF in quotes): pandas decides
        c->row_begin.push_back(line_begin(k));
        c->row_end.push_back(line_end(k));
    }
    if (c->n_cols <= 0) c->flags |= FLAG_UNSUPPORTED;
    return 0;
}

extern "C" int dyd_csv_info(void* handle, int64_t* n_rows, int32_t* n_cols, int64_t* header_begin, int64_t* header_end,
                            int32_t* flags) {
    if (!handle) return DYD_E_ARG;
    const Csv* c = (const Csv*)handle;
    if (n_rows) *n_rows = (int64_t)c->row_begin.size();
    if (n_cols) *n_cols = c->n_cols;
    if (header_begin) *header_begin = c->header_begin;
    if (header_end) *header_end = c->header_end;
    if (flags) *flags = c->flags;
    return 0;
}

// Sizes every cell.  col_bytes[j] = unescaped bytes of column j, col_nulls[j] = missing cells,
// col_text[j] = 1 iff every window of `window` consecutive rows holds a cell that is certainly text
// (pandas infers dtypes per chunk of that many rows), col_utf8[j] = 1 iff all cells are valid UTF-8.
extern "C" int dyd_csv_measure(void* handle, int64_t window, int64_t* col_bytes, int64_t* col_nulls, uint8_t* col_text,
                               uint8_t* col_utf8, int32_t threads) {
    if (!handle || window <= 0 || !col_bytes || !col_nulls || !col_text || !col_utf8) return DYD_E_ARG;
    Csv* c = (Csv*)handle;
    if (c->flags & FLAG_UNSUPPORTED) return DYD_E_ARG;
    const int64_t nr = (int64_t)c->row_begin.size();
    const int nc = c->n_cols;
    c->off.assign(nc, std::vector<int64_t>((size_t)nr + 1, 0));
    c->isna.assign(nc, std::vector<uint8_t>((size_t)nr, 0));
    const int64_t n_win = std::max<int64_t>((nr + window - 1) / window, 1);
    int T = threads > 0 ? threads : (int)std::max(1u, std::thread::hardware_concurrency());
    // per thread: one flag per (column, window) and per column, merged after the join
    std::vector<std::vector<uint8_t>> t_text((size_t)T), t_bad((size_t)T);
    const uint8_t* d = c->d;
    if (c->fast) {
        parallel_ranges(nr, T, 2048, [&](int64_t ra, int64_t rb, int t) {
            auto& text = t_text[t];
            auto& bad = t_bad[t];
            text.assign((size_t)(nc * n_win), 0);
            bad.assign((size_t)nc, 0);
            uint8_t small[32];
            for (int64_t r = ra; r < rb; ++r) {
                const int64_t w = r / window;
                int64_t b = c->row_begin[r];
                for (int f = 0; f < nc; ++f) {
                    const int64_t e = c->fend[(size_t)r * nc + f];
                    const uint32_t nq = c->fquotes[(size_t)r * nc + f];
                    const uint8_t* s = d + b;
                    int64_t len = e - b;
                    if (nq) { ++s; len = len - 2 - (int64_t)(nq - 2) / 2; }       // quoted: drop the outer quotes, "" counts once
                    c->off[f][r + 1] = len;
                    const uint8_t* content = s;
                    if (nq > 2 && len <= 9) { unescape_scalar(small, s, e - b - 2); content = small; }
                    if (c->is_na(content, len)) c->isna[f][r] = 1;
                    else if (certainly_text(content, len)) text[(size_t)(f * n_win + w)] = 1;
                    if (len && fast_any_high_bit(d + b, e - b) && !utf8_ok(d + b, e - b)) bad[f] = 1;
                    b = e + 1;
                }
            }
        });
    } else
    parallel_ranges(nr, T, 2048, [&](int64_t ra, int64_t rb, int t) {
        auto& text = t_text[t];
        auto& bad = t_bad[t];
        text.assign((size_t)(nc * n_win), 0);
        bad.assign((size_t)nc, 0);
        std::string scratch;
        for (int64_t r = ra; r < rb; ++r) {
            const int64_t w = r / window;
            int pieces = 0;
            int64_t len = 0;
            const uint8_t* first = nullptr;
            walk_record(d + c->row_begin[r], d + c->row_end[r],
                [&](int, const uint8_t* p, int64_t l) {
                    if (pieces == 0) first = p;
                    else { if (pieces == 1) scratch.assign((const char*)first, (size_t)len); scratch.append((const char*)p, (size_t)l); }
                    ++pieces; len += l;
                },
                [&](int f) {
                    const uint8_t* s = pieces <= 1 ? first : (const uint8_t*)scratch.data();
                    c->off[f][r + 1] = len;
                    if (c->is_na(s, len)) c->isna[f][r] = 1;
                    else if (certainly_text(s, len)) text[(size_t)(f * n_win + w)] = 1;
                    if (len && !utf8_ok(s, len)) bad[f] = 1;
                    pieces = 0; len = 0; first = nullptr;
                });
        }
    });
    std::vector<std::vector<uint8_t>> win_text(nc, std::vector<uint8_t>((size_t)n_win, 0));
    std::vector<uint8_t> bad_utf8((size_t)nc, 0);
    for (int t = 0; t < T; ++t) {
        if (t_text[t].empty()) continue;
        for (int j = 0; j < nc; ++j) {
            bad_utf8[j] |= t_bad[t][j];
            for (int64_t w = 0; w < n_win; ++w) win_text[j][w] |= t_text[t][(size_t)(j * n_win + w)];
        }
    }
    for (int j = 0; j < nc; ++j) {
        int64_t acc = 0, nulls = 0;
        auto& o = c->off[j];
        for (int64_t r = 0; r < nr; ++r) { const int64_t l = o[r + 1]; o[r] = acc; acc += c->isna[j][r] ? 0 : l; nulls += c->isna[j][r]; }
        o[nr] = acc;
        col_bytes[j] = acc; col_nulls[j] = nulls;
        uint8_t all = nr > 0;
        for (int64_t w = 0; w < (nr + window - 1) / window; ++w) all &= win_text[j][w];
        col_text[j] = all; col_utf8[j] = !bad_utf8[j];
    }
    return 0;
}

// Writes the selected columns: Arrow large_string buffers (offsets int64[n_rows+1], data, validity
// bitmap (n_rows+7)/8 bytes, LSB first; may be NULL when the column has no missing cell).
extern "C" int dyd_csv_fill(void* handle, int32_t n_sel, const int32_t* cols, int64_t* const* off_out, uint8_t* const* data_out,
                            uint8_t* const* bitmap_out, int32_t threads) {
    if (!handle || n_sel < 0 || (n_sel > 0 && (!cols || !off_out || !data_out || !bitmap_out))) return DYD_E_ARG;
    Csv* c = (Csv*)handle;
    if (c->off.empty() && c->n_cols > 0) return DYD_E_ARG;
    const int64_t nr = (int64_t)c->row_begin.size();
    std::vector<int> slot((size_t)c->n_cols, -1);
    for (int k = 0; k < n_sel; ++k) {
        if (cols[k] < 0 || cols[k] >= c->n_cols || !off_out[k] || !data_out[k]) return DYD_E_ARG;
        slot[cols[k]] = k;
        memcpy(off_out[k], c->off[cols[k]].data(), sizeof(int64_t) * (size_t)(nr + 1));
    }
    const uint8_t* d = c->d;
    if (c->fast) {
        const int nc = c->n_cols;
        parallel_ranges(nr, threads, 512, [&](int64_t a, int64_t b, int) {
            for (int64_t r = a; r < b; ++r) {
                int64_t fb = c->row_begin[r];
                for (int f = 0; f < nc; ++f) {
                    const int64_t e = c->fend[(size_t)r * nc + f];
                    const int k = slot[f];
                    if (k >= 0 && !c->isna[f][r]) {
                        const uint32_t nq = c->fquotes[(size_t)r * nc + f];
                        uint8_t* o = data_out[k] + c->off[f][r];
                        if (nq <= 2) memcpy(o, d + fb + (nq ? 1 : 0), (size_t)(e - fb - (nq ? 2 : 0)));
                        else fast_unescape(o, d + fb + 1, e - fb - 2);
                    }
                    fb = e + 1;
                }
            }
        });
    } else
    parallel_ranges(nr, threads, 2048, [&](int64_t a, int64_t b, int) {
        for (int64_t r = a; r < b; ++r) {
            int64_t pos = 0;
            int cur = -1;
            walk_record(d + c->row_begin[r], d + c->row_end[r],
                [&](int f, const uint8_t* p, int64_t l) {
                    const int k = slot[f];
                    if (k < 0 || c->isna[f][r]) return;
                    if (cur != f) { cur = f; pos = c->off[f][r]; }
                    memcpy(data_out[k] + pos, p, (size_t)l);
                    pos += l;
                },
                [](int) {});
        }
    });
    for (int k = 0; k < n_sel; ++k) {
        if (!bitmap_out[k]) continue;
        const auto& na = c->isna[cols[k]];
        uint8_t* bm = bitmap_out[k];
        parallel_ranges((nr + 7) / 8, threads, 1 << 16, [&](int64_t a, int64_t b, int) {
            for (int64_t i = a; i < b; ++i) {
                uint8_t v = 0;
                for (int bit = 0; bit < 8 && i * 8 + bit < nr; ++bit) v |= (uint8_t)((na[i * 8 + bit] ? 0 : 1) << bit);
                bm[i] = v;
            }
        });
    }
    return 0;
}

extern "C" void dyd_csv_close(void* handle) { delete (Csv*)handle; }

// The first n bytes of a file into `out`, several threads each reading its own range with pread(2) (the copy out of
// the page cache and the first touch of `out` both spread over the cores).  DYD_E_IO (errno set) on failure / short file.
extern "C" int dyd_read_file(const char* path, uint8_t* out, int64_t n, int32_t threads) {
    if (!path || n < 0 || (n > 0 && !out)) return DYD_E_ARG;
    const int fd = ::open(path, O_RDONLY | O_CLOEXEC);
    if (fd < 0) return DYD_E_IO;
    std::atomic<int> err{0};
    parallel_ranges(n, threads, 4 << 20, [&](int64_t a, int64_t b, int) {
        while (a < b && !err.load(std::memory_order_relaxed)) {
            const ssize_t r = ::pread(fd, out + a, (size_t)std::min<int64_t>(b - a, 8 << 20), (off_t)a);
            if (r < 0) { if (errno == EINTR) continue; err.store(errno); return; }
            if (r == 0) { err.store(EIO); return; }               // the file is shorter than the caller was told
            a += r;
        }
    });
    ::close(fd);
    if (err.load()) { errno = err.load(); return DYD_E_IO; }
    return 0;
}

// Would pd.read_csv give back this text column after DataFrame.to_csv wrote it?  (deal_yolo_daya_b200/tablecache.py keeps
// the frame a step wrote so that the next step does not parse the file again; that is only sound for columns that survive
// the round trip.)  Yes iff no valid cell is empty or one of pandas' NA strings (it would come back as NaN), no cell holds
// a NUL byte, and every dtype-inference window of `window` rows holds a cell that is certainly text (the column then stays
// a string column in every chunk).  valid = one byte per row (NULL: all valid); check_cells = 0 skips the per-cell part
// for columns whose cells are known to come from this reader.  Returns 1 / 0, or DYD_E_ARG.
extern "C" int dyd_csv_roundtrip_check(const int64_t* off, const uint8_t* data, const uint8_t* valid, int64_t n_rows, int64_t window,
                                       const uint8_t* na_bytes, const int64_t* na_off, int32_t n_na, int32_t check_cells, int32_t threads) {
    if (!off || n_rows < 0 || window <= 0 || (n_rows > 0 && !data && off[n_rows] > off[0]) || n_na < 0 || (n_na > 0 && (!na_bytes || !na_off))) return DYD_E_ARG;
    if (n_rows == 0) return 0;
    Csv c;
    for (int32_t k = 0; k < n_na; ++k) c.na.emplace_back((const char*)na_bytes + na_off[k], (size_t)(na_off[k + 1] - na_off[k]));
    const int64_t n_win = (n_rows + window - 1) / window;
    int T = threads > 0 ? threads : (int)std::max(1u, std::thread::hardware_concurrency());
    std::vector<std::vector<uint8_t>> t_win((size_t)T);
    std::atomic<int> bad{0};
    parallel_ranges(n_rows, T, 4096, [&](int64_t ra, int64_t rb, int t) {
        auto& win = t_win[(size_t)t];
        win.assign((size_t)n_win, 0);
        for (int64_t r = ra; r < rb; ++r) {
            if (valid && !valid[r]) continue;
            const uint8_t* s = data + off[r];
            const int64_t len = off[r + 1] - off[r];
            uint8_t& text = win[(size_t)(r / window)];
            if (!text && certainly_text(s, len)) text = 1;
            if (check_cells) {
                if (c.is_na(s, len) || memchr(s, 0, (size_t)len)) { bad.store(1); return; }
            } else if (text) {
                r = std::min(rb, (r / window + 1) * window) - 1;        // this window is settled: on to the next one
            }
        }
    });
    std::vector<uint8_t> win_ok((size_t)n_win, 0);
    for (const auto& win : t_win) for (size_t w = 0; w < win.size(); ++w) win_ok[w] |= win[w];
    if (bad.load()) return 0;
    for (int64_t w = 0; w < n_win; ++w) if (!win_ok[(size_t)w]) return 0;
    return 1;
}

// Native columnar ingest / egress of the annotation JSON cells, host C++.
//
// ingest  : UTF-8 cell texts -> ragged CSR (img_off, poly_off, xy) for step 4 (polygons) or
//           (img_off, pts, valid) for step 5 (two-point boxes), multi-threaded, one pass per row.
// egress  : step 4's output cell = the reference's json.dumps(ensure_ascii=False) of the parsed
//           document with every ptList replaced by two corner points (processor.py:262-279).  When the
//           input text is already in json.dumps' canonical form (which it is whenever a json.dumps-
//           based tool wrote it) that output equals the input text with each ptList value spliced
//           out for `[{"x": X1, "y": Y1}, {"x": X2, "y": Y2}]`, X/Y being the ORIGINAL number literals
//           of the vertices K1 selected -- no float formatting, no Python objects.
//
// Anything this file is not certain about is NOT guessed: the row is flagged SLOW and the Python
// layer (ingest.py) handles it with CPython's own json module.  SLOW covers: text that is not
// canonical (strict mode), non-dict documents / elements, duplicate keys, values an fp64 array
// cannot carry (null / strings / ints beyond 2^53 as coordinates), malformed JSON.
// Canonical form is verified, not assumed: separators, string escapes, and every number literal
// must equal CPython's repr of its parsed value (shortest round-trip digits, repr's fixed/exponent
// switch at 1e16 / 1e-4, ".0" suffix), computed here with std::to_chars / std::from_chars.
#include <fcntl.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <charconv>
#include <condition_variable>
#include <mutex>
#include <cmath>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <string>
#include <string_view>
#include <thread>
#include <vector>

#include "../../include/dyd.h"
#include "simd_text.hpp"

namespace {

enum RowStatus : uint8_t { ROW_OK = 0, ROW_SLOW = 1, ROW_NOT_TEXT = 2 };
enum Kind : uint8_t { K_OBJ, K_ARR, K_STR, K_INT, K_FLT, K_TRUE, K_FALSE, K_NULL };

struct Span { uint32_t off, len; };

struct Vertex { double x, y; Span sx, sy; };

struct Poly {
    uint32_t splice_off, splice_len;     // bytes of the row text replaced by the new ptList value
    uint8_t splice_kind;                 // 0: replace ptList value; 1: append `"ptList": V` into polygon dict;
                                         // 2: append `"polygon": {"ptList": V}` into the object
    uint8_t container_empty;             // for kinds 1/2: the dict being appended to is `{}`
    uint32_t v_begin, v_count;           // vertices (row-local index into RowOut::verts)
};

struct RowOut {
    uint8_t status = ROW_SLOW;
    std::vector<Poly> polys;
    std::vector<Vertex> verts;
    Span width{0, 0}, height{0, 0};      // literal spans of top-level width / height (len 0 = absent)
    uint8_t width_kind = K_NULL, height_kind = K_NULL;
    // step 5
    std::vector<double> boxes;           // 4 per box
    std::vector<uint8_t> bvalid;
    // step 5.5 / 6 (mode 2): per dict object the span of its "name" string contents (len -1: None / absent)
    // and of the object itself; the "objects" member (key .. value) and how it sits among its siblings
    std::vector<Span> names, objspans;
    std::string canon;                   // json.dumps(json.loads(text)) when the text was valid JSON in another style:
                                         // every span of this row then refers to `canon`, not to the input text
    Span member{0, 0};
    uint8_t member_first = 0, member_has_next = 0;
    int32_t list_len = 0;                // elements of the "objects" list, dicts or not
};
enum RowStatus2 : uint8_t { ROW_NO_LIST = 4, ROW_NOT_A_LIST = 5 };   // mode 2: "objects" absent / present but not a list

struct Fail {};

// keys of one JSON object (raw spans), inline storage: no heap traffic per dict
struct KeySet {
    static constexpr int CAP = 32;
    Span k[CAP];
    int n = 0;
};

// ---------------------------------------------------------------- CPython float repr
// repr(float) from the shortest round-trip digits (float_repr_style "short", format code 'r').
static int py_float_repr(double v, char* out) {
    if (std::isnan(v)) { memcpy(out, "NaN", 3); return 3; }
    if (std::isinf(v)) { if (v > 0) { memcpy(out, "Infinity", 8); return 8; } memcpy(out, "-Infinity", 9); return 9; }
    char* p = out;
    if (std::signbit(v)) { *p++ = '-'; v = -v; }
    if (v == 0.0) { memcpy(p, "0.0", 3); return (int)(p - out) + 3; }
    char sci[40];
    auto r = std::to_chars(sci, sci + sizeof(sci), v, std::chars_format::scientific);   // shortest round-trip
    // sci = d[.ddd]e[+-]XX
    char digits[24]; int nd = 0;
    const char* q = sci;
    for (; q < r.ptr && *q != 'e'; ++q) if (*q != '.') digits[nd++] = *q;
    int e10 = 0; { ++q; bool neg = *q == '-'; ++q; for (; q < r.ptr; ++q) e10 = e10 * 10 + (*q - '0'); if (neg) e10 = -e10; }
    const int decpt = e10 + 1;
    if (decpt <= -4 || decpt > 16) {
        *p++ = digits[0];
        if (nd > 1) { *p++ = '.'; memcpy(p, digits + 1, nd - 1); p += nd - 1; }
        *p++ = 'e';
        int ex = decpt - 1;
        *p++ = ex < 0 ? '-' : '+';
        if (ex < 0) ex = -ex;
        char eb[8]; int ne = 0;
        do { eb[ne++] = (char)('0' + ex % 10); ex /= 10; } while (ex);
        if (ne < 2) eb[ne++] = '0';
        while (ne) *p++ = eb[--ne];
    } else if (decpt <= 0) {
        *p++ = '0'; *p++ = '.';
        for (int i = 0; i < -decpt; ++i) *p++ = '0';
        memcpy(p, digits, nd); p += nd;
    } else if (decpt >= nd) {
        memcpy(p, digits, nd); p += nd;
        for (int i = 0; i < decpt - nd; ++i) *p++ = '0';
        *p++ = '.'; *p++ = '0';
    } else {
        memcpy(p, digits, decpt); p += decpt;
        *p++ = '.';
        memcpy(p, digits + decpt, nd - decpt); p += nd - decpt;
    }
    return (int)(p - out);
}

// ---------------------------------------------------------------- tokenizer
struct Parser {
    const char* base;       // row text
    const char* p;
    const char* end;
    bool strict;            // canonical json.dumps(ensure_ascii=False) formatting required

    [[noreturn]] static void fail() { throw Fail{}; }
    void need(bool c) const { if (!c) fail(); }
    uint32_t at() const { return (uint32_t)(p - base); }

    void ws() { if (!strict) while (p < end && (*p == ' ' || *p == '\n' || *p == '\r' || *p == '\t')) ++p; }
    void lit(const char* s, size_t n) { need((size_t)(end - p) >= n && memcmp(p, s, n) == 0); p += n; }
    // after an item inside a container: returns true if another item follows
    bool more(char close) {
        ws();
        need(p < end);
        if (*p == close) { ++p; return false; }
        need(*p == ','); ++p;
        if (strict) { need(p < end && *p == ' '); ++p; } else ws();
        return true;
    }
    void colon() {
        ws();
        need(p < end && *p == ':'); ++p;
        if (strict) { need(p < end && *p == ' '); ++p; } else ws();
    }
    // string: returns the span of the raw contents (between the quotes); `plain` = no escapes inside
    Span str(bool* plain = nullptr) {
        need(p < end && *p == '"'); ++p;
        const char* s = p;
        bool pl = true;
        for (;;) {
            need(p < end);
            unsigned char c = (unsigned char)*p;
            if (c == '"') break;
            if (c == '\\') {
                pl = false;
                need(p + 1 < end);
                char e = p[1];
                if (e == 'u') {
                    need(p + 5 < end);
                    unsigned cp = 0;
                    for (int i = 2; i < 6; ++i) {
                        char h = p[i]; unsigned d;
                        if (h >= '0' && h <= '9') d = h - '0';
                        else if (h >= 'a' && h <= 'f') d = h - 'a' + 10;
                        else if (h >= 'A' && h <= 'F') { d = h - 'A' + 10; if (strict) fail(); }
                        else fail();
                        cp = cp * 16 + d;
                    }
                    // json.dumps(ensure_ascii=False) writes \u only for control characters without a short escape
                    if (strict) need(cp < 0x20 && cp != '\n' && cp != '\r' && cp != '\t' && cp != '\b' && cp != '\f');
                    p += 6;
                } else {
                    bool ok = e == '"' || e == '\\' || e == 'n' || e == 'r' || e == 't' || e == 'b' || e == 'f';
                    if (!strict) ok = ok || e == '/';
                    need(ok);
                    p += 2;
                }
                continue;
            }
            need(c >= 0x20);                // raw control characters are not valid JSON (json.loads strict=True)
            ++p;
        }
        Span sp{(uint32_t)(s - base), (uint32_t)(p - s)};
        ++p;
        if (plain) *plain = pl;
        return sp;
    }
    // number / literal scalar: returns kind, literal span, value (for numbers and bools)
    Kind scalar(Span* sp, double* val, bool* exact = nullptr) {
        need(p < end);
        if (exact) *exact = true;
        const char* s = p;
        char c = *p;
        if (c == 't') { lit("true", 4); *sp = {(uint32_t)(s - base), 4}; *val = 1.0; return K_TRUE; }
        if (c == 'f') { lit("false", 5); *sp = {(uint32_t)(s - base), 5}; *val = 0.0; return K_FALSE; }
        if (c == 'n') { lit("null", 4); *sp = {(uint32_t)(s - base), 4}; *val = 0.0; return K_NULL; }
        if (c == 'N') { lit("NaN", 3); *sp = {(uint32_t)(s - base), 3}; *val = std::nan(""); return K_FLT; }
        if (c == 'I') { lit("Infinity", 8); *sp = {(uint32_t)(s - base), 8}; *val = INFINITY; return K_FLT; }
        if (c == '-' && p + 1 < end && p[1] == 'I') { lit("-Infinity", 9); *sp = {(uint32_t)(s - base), 9}; *val = -INFINITY; return K_FLT; }
        // JSON number grammar: -?(0|[1-9]\d*)(\.\d+)?([eE][+-]?\d+)?
        if (*p == '-') ++p;
        need(p < end && *p >= '0' && *p <= '9');
        if (*p == '0') ++p; else while (p < end && *p >= '0' && *p <= '9') ++p;
        bool is_float = false;
        if (p < end && *p == '.') { is_float = true; ++p; need(p < end && *p >= '0' && *p <= '9'); while (p < end && *p >= '0' && *p <= '9') ++p; }
        if (p < end && (*p == 'e' || *p == 'E')) {
            is_float = true; ++p;
            if (p < end && (*p == '+' || *p == '-')) ++p;
            need(p < end && *p >= '0' && *p <= '9');
            while (p < end && *p >= '0' && *p <= '9') ++p;
        }
        *sp = {(uint32_t)(s - base), (uint32_t)(p - s)};
        if (is_float) {
            double v = 0.0;
            auto r = std::from_chars(s, p, v);            // correctly rounded, like float()
            if (r.ec == std::errc::result_out_of_range) {
                // float("1e999") is inf, float("1e-999") is 0.0: redo with strtod semantics
                std::string tmp(s, p); v = strtod(tmp.c_str(), nullptr);
            } else need(r.ec == std::errc() && r.ptr == p);
            *val = v;
            if (strict) {                                   // literal must be repr(v)
                char buf[40]; int n = py_float_repr(v, buf);
                need((size_t)n == (size_t)(p - s) && memcmp(buf, s, n) == 0);
            }
            return K_FLT;
        }
        // integer: exact in fp64 up to 2^53 (larger ones may be carried along, not used as coordinates);
        // "-0" is not what json.dumps writes for an int
        const char* d = s; bool neg = false;
        if (*d == '-') { neg = true; ++d; }
        uint64_t u = 0; bool ok = (size_t)(p - d) <= 16;
        if (ok) { for (const char* t = d; t < p; ++t) u = u * 10 + (uint64_t)(*t - '0'); ok = u <= (1ULL << 53); }
        if (strict && neg) need(!(ok && u == 0));
        if (exact) *exact = ok;
        *val = ok ? (neg ? -(double)u : (double)u) : 0.0;
        return K_INT;
    }
    // skip any value (validating it); returns its kind
    Kind skip() {
        need(p < end);
        if (*p == '{') {
            ++p; ws();
            if (p < end && *p == '}') { ++p; return K_OBJ; }
            KeySet keys;
            do { Span k = str(); check_dup(keys, k); colon(); skip(); } while (more('}'));
            return K_OBJ;
        }
        if (*p == '[') {
            ++p; ws();
            if (p < end && *p == ']') { ++p; return K_ARR; }
            do { skip(); } while (more(']'));
            return K_ARR;
        }
        if (*p == '"') { str(); return K_STR; }
        Span sp; double v;
        Kind k = scalar(&sp, &v);
        // a huge integer literal is fine when it is only carried along, scalar() already bounded it;
        return k;
    }
    // json.loads keeps the last of duplicate keys, json.dumps then writes one: a splice would differ
    void check_dup(KeySet& keys, Span k) {
        for (int i = 0; i < keys.n; ++i) if (keys.k[i].len == k.len && memcmp(base + keys.k[i].off, base + k.off, k.len) == 0) fail();
        need(keys.n < KeySet::CAP);          // very wide dicts go to the slow lane
        keys.k[keys.n++] = k;
    }
    bool key_is(Span k, const char* s) const { size_t n = strlen(s); return k.len == n && memcmp(base + k.off, s, n) == 0; }
};

// ---------------------------------------------------------------- canonical form
// json.dumps(json.loads(text), ensure_ascii=False) without CPython: any valid JSON text (other separators,
// indentation, \u escapes, "1.50", "1E5" ...) is rewritten in the one style the splicing lanes work on.
// The parsed VALUE is untouched by construction (same ints, same floats, same strings), so whatever the
// reference computes from json.loads(text) it also computes from json.loads(canonical text).  Documents
// this file is not certain about (duplicate keys, escaped keys, lone surrogates, nesting deeper than 200,
// anything json.loads rejects) fail and stay with the CPython lane.
static void canon_string(Parser& ps, std::string& o, bool is_key) {
    ps.need(ps.p < ps.end && *ps.p == '"'); ++ps.p;
    o.push_back('"');
    auto put_cp = [&](unsigned cp) {
        switch (cp) {
            case '"': o += "\\\""; return;
            case '\\': o += "\\\\"; return;
            case '\n': o += "\\n"; return;
            case '\r': o += "\\r"; return;
            case '\t': o += "\\t"; return;
            case '\b': o += "\\b"; return;
            case '\f': o += "\\f"; return;
            default: break;
        }
        if (cp < 0x20) { char b[8]; snprintf(b, sizeof b, "\\u%04x", cp); o += b; return; }
        if (cp < 0x80) o.push_back((char)cp);
        else if (cp < 0x800) { o.push_back((char)(0xC0 | (cp >> 6))); o.push_back((char)(0x80 | (cp & 0x3F))); }
        else if (cp < 0x10000) { o.push_back((char)(0xE0 | (cp >> 12))); o.push_back((char)(0x80 | ((cp >> 6) & 0x3F))); o.push_back((char)(0x80 | (cp & 0x3F))); }
        else { o.push_back((char)(0xF0 | (cp >> 18))); o.push_back((char)(0x80 | ((cp >> 12) & 0x3F))); o.push_back((char)(0x80 | ((cp >> 6) & 0x3F))); o.push_back((char)(0x80 | (cp & 0x3F))); }
    };
    auto hex4 = [&](const char* q) {
        unsigned cp = 0;
        for (int i = 0; i < 4; ++i) {
            const char h = q[i]; unsigned d;
            if (h >= '0' && h <= '9') d = h - '0';
            else if (h >= 'a' && h <= 'f') d = h - 'a' + 10;
            else if (h >= 'A' && h <= 'F') d = h - 'A' + 10;
            else Parser::fail();
            cp = cp * 16 + d;
        }
        return cp;
    };
    for (;;) {
        ps.need(ps.p < ps.end);
        const unsigned char c = (unsigned char)*ps.p;
        if (c == '"') { ++ps.p; break; }
        if (c == '\\') {
            ps.need(ps.p + 1 < ps.end);
            const char e = ps.p[1];
            if (e == 'u') {
                ps.need(ps.p + 5 < ps.end);
                unsigned cp = hex4(ps.p + 2);
                ps.p += 6;
                if (cp >= 0xD800 && cp <= 0xDBFF) {
                    ps.need(ps.p + 5 < ps.end && ps.p[0] == '\\' && ps.p[1] == 'u');
                    const unsigned lo = hex4(ps.p + 2);
                    ps.need(lo >= 0xDC00 && lo <= 0xDFFF);               // a lone surrogate cannot be written as UTF-8
                    ps.p += 6;
                    cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                } else ps.need(!(cp >= 0xDC00 && cp <= 0xDFFF));
                put_cp(cp);
            } else {
                unsigned cp;
                switch (e) {
                    case '"': cp = '"'; break;  case '\\': cp = '\\'; break;  case '/': cp = '/'; break;
                    case 'b': cp = '\b'; break; case 'f': cp = '\f'; break; case 'n': cp = '\n'; break;
                    case 'r': cp = '\r'; break; case 't': cp = '\t'; break;
                    default: Parser::fail();
                }
                ps.p += 2;
                put_cp(cp);
            }
            continue;
        }
        ps.need(c >= 0x20);                         // raw control characters are not valid JSON
        o.push_back((char)c);
        ++ps.p;
    }
    o.push_back('"');
    (void)is_key;
}

static void canon_value(Parser& ps, std::string& o, int depth) {
    ps.need(depth < 200);
    ps.ws();
    ps.need(ps.p < ps.end);
    const char c = *ps.p;
    if (c == '{') {
        ++ps.p; ps.ws();
        o.push_back('{');
        if (ps.p < ps.end && *ps.p == '}') { ++ps.p; o.push_back('}'); return; }
        // duplicate keys (json.loads keeps the last value at the first position) are left to CPython; keys are
        // compared in their canonical spelling, which is unique per string
        size_t key_at[KeySet::CAP], key_len[KeySet::CAP];
        int n_keys = 0;
        for (;;) {
            ps.ws();
            const size_t o_at = o.size();
            canon_string(ps, o, true);
            const size_t kl = o.size() - o_at;
            for (int i = 0; i < n_keys; ++i) ps.need(!(key_len[i] == kl && memcmp(o.data() + key_at[i], o.data() + o_at, kl) == 0));
            ps.need(n_keys < KeySet::CAP);
            key_at[n_keys] = o_at; key_len[n_keys] = kl; ++n_keys;
            ps.colon();
            o += ": ";
            canon_value(ps, o, depth + 1);
            if (!ps.more('}')) break;
            o += ", ";
        }
        o.push_back('}');
        return;
    }
    if (c == '[') {
        ++ps.p; ps.ws();
        o.push_back('[');
        if (ps.p < ps.end && *ps.p == ']') { ++ps.p; o.push_back(']'); return; }
        for (;;) {
            canon_value(ps, o, depth + 1);
            if (!ps.more(']')) break;
            o += ", ";
        }
        o.push_back(']');
        return;
    }
    if (c == '"') { canon_string(ps, o, false); return; }
    Span sp; double v;
    const Kind k = ps.scalar(&sp, &v);
    switch (k) {
        case K_TRUE: o += "true"; break;
        case K_FALSE: o += "false"; break;
        case K_NULL: o += "null"; break;
        case K_FLT: { char b[40]; const int n = py_float_repr(v, b); o.append(b, (size_t)n); break; }
        default: {                                  // int: the digits as written; int("-0") is 0
            const char* q = ps.base + sp.off;
            if (sp.len >= 2 && q[0] == '-' && q[1] == '0') o.push_back('0'); else o.append(q, sp.len);
        }
    }
}

// whole document; fails (throws) unless the text is exactly one JSON value surrounded by whitespace
static void canonicalize(const char* text, size_t len, std::string& o) {
    Parser ps{text, text, text + len, false};
    o.clear();
    o.reserve(len + len / 8);
    canon_value(ps, o, 0);
    ps.ws();
    ps.need(ps.p == ps.end);
}

// ---------------------------------------------------------------- step 4: polygons
static void parse_point(Parser& ps, RowOut& out, bool& valid_point) {
    // one element of a ptList; a valid point is a dict with both "x" and "y" (processor.py:253)
    valid_point = false;
    ps.need(ps.p < ps.end);
    if (*ps.p != '{') { ps.skip(); return; }
    ++ps.p; ps.ws();
    if (ps.p < ps.end && *ps.p == '}') { ++ps.p; return; }
    KeySet keys;
    bool hx = false, hy = false, num_ok = true;
    Vertex v{};
    do {
        bool plain; Span k = ps.str(&plain); ps.check_dup(keys, k); ps.colon();
        bool isx = plain && ps.key_is(k, "x"), isy = plain && ps.key_is(k, "y");
        if (isx || isy) {
            ps.need(ps.p < ps.end);
            if (*ps.p == '{' || *ps.p == '[' || *ps.p == '"') { ps.skip(); num_ok = false; }
            else {
                Span sp; double val; bool exact; Kind kd = ps.scalar(&sp, &val, &exact);
                if (kd == K_NULL || !exact) num_ok = false;
                if (isx) { hx = true; v.x = val; v.sx = sp; } else { hy = true; v.y = val; v.sy = sp; }
            }
        } else ps.skip();
    } while (ps.more('}'));
    if (hx && hy) {
        if (!num_ok) Parser::fail();          // None / string / container coordinate: CPython semantics -> slow lane
        valid_point = true;
        out.verts.push_back(v);
    }
}

static void parse_polygon_row(const char* text, size_t len, bool strict, RowOut& out) {
    Parser ps{text, text, text + len, strict};
    ps.ws();
    ps.need(ps.p < ps.end && *ps.p == '{');
    ++ps.p; ps.ws();
    bool have_objects = false;
    KeySet top_keys;
    if (ps.p < ps.end && *ps.p == '}') { ++ps.p; }
    else do {
        bool plain; Span k = ps.str(&plain); ps.check_dup(top_keys, k); ps.colon();
        if (plain && ps.key_is(k, "objects")) {
            have_objects = true;
            ps.need(ps.p < ps.end && *ps.p == '[');          // a non-list "objects" is iterated differently by the reference
            ++ps.p; ps.ws();
            if (ps.p < ps.end && *ps.p == ']') { ++ps.p; continue; }
            do {
                ps.need(ps.p < ps.end && *ps.p == '{');      // non-dict objects are dropped by the reference: no splice
                const uint32_t obj_open = ps.at();
                ++ps.p; ps.ws();
                Poly poly{}; poly.v_begin = (uint32_t)out.verts.size();
                bool have_polygon = false;
                bool obj_empty = false;
                KeySet okeys;
                if (ps.p < ps.end && *ps.p == '}') { obj_empty = true; }
                else do {
                    bool pl2; Span ok = ps.str(&pl2); ps.check_dup(okeys, ok); ps.colon();
                    if (pl2 && ps.key_is(ok, "polygon")) {
                        have_polygon = true;
                        ps.need(ps.p < ps.end && *ps.p == '{');          // .get on a non-dict raises in the reference
                        ++ps.p; ps.ws();
                        bool have_pt = false, pg_empty = false;
                        KeySet pkeys;
                        if (ps.p < ps.end && *ps.p == '}') { pg_empty = true; }
                        else do {
                            bool pl3; Span pk = ps.str(&pl3); ps.check_dup(pkeys, pk); ps.colon();
                            if (pl3 && ps.key_is(pk, "ptList")) {
                                have_pt = true;
                                ps.need(ps.p < ps.end && *ps.p == '[');  // iterating a non-list ptList: slow lane
                                const uint32_t s = ps.at();
                                ++ps.p; ps.ws();
                                if (ps.p < ps.end && *ps.p == ']') { ++ps.p; }
                                else do { bool vp; parse_point(ps, out, vp); } while (ps.more(']'));
                                poly.splice_off = s; poly.splice_len = ps.at() - s; poly.splice_kind = 0;
                            } else ps.skip();
                        } while (ps.more('}'));
                        if (pg_empty) ++ps.p;
                        if (!have_pt) {                                  // "ptList" appended to the polygon dict
                            poly.splice_off = ps.at() - 1; poly.splice_len = 0; poly.splice_kind = 1; poly.container_empty = pg_empty;
                        }
                    } else ps.skip();
                } while (ps.more('}'));
                if (obj_empty) ++ps.p;
                if (!have_polygon) {                                     // "polygon" appended to the object
                    poly.splice_off = ps.at() - 1; poly.splice_len = 0; poly.splice_kind = 2; poly.container_empty = obj_empty;
                }
                (void)obj_open;
                poly.v_count = (uint32_t)out.verts.size() - poly.v_begin;
                out.polys.push_back(poly);
            } while (ps.more(']'));
        } else if (plain && (ps.key_is(k, "width") || ps.key_is(k, "height"))) {
            ps.need(ps.p < ps.end);
            const bool w = ps.key_is(k, "width");
            if (*ps.p == '{' || *ps.p == '[' || *ps.p == '"') Parser::fail();   // Python object needed: slow lane
            Span sp; double val; Kind kd = ps.scalar(&sp, &val);
            if (w) { out.width = sp; out.width_kind = kd; } else { out.height = sp; out.height_kind = kd; }
        } else ps.skip();
    } while (ps.more('}'));
    ps.ws();
    ps.need(ps.p == ps.end);
    if (!have_objects) Parser::fail();        // the reference appends "objects": [] -- left to the slow lane
    out.status = ROW_OK;
}

// ---------------------------------------------------------------- step 5: two-point boxes (lenient)
static void parse_box_row(const char* text, size_t len, RowOut& out) {
    Parser ps{text, text, text + len, false};
    ps.ws();
    ps.need(ps.p < ps.end && *ps.p == '{');
    ++ps.p; ps.ws();
    KeySet top_keys;
    bool stop = false;                       // an object raised in the reference: later objects are not looked at
    if (ps.p < ps.end && *ps.p == '}') { ++ps.p; }
    else do {
        bool plain; Span k = ps.str(&plain); ps.check_dup(top_keys, k); ps.colon();
        if (plain && ps.key_is(k, "objects")) {
            ps.need(ps.p < ps.end && *ps.p == '[');
            ++ps.p; ps.ws();
            if (ps.p < ps.end && *ps.p == ']') { ++ps.p; continue; }
            do {
                ps.need(ps.p < ps.end);
                if (*ps.p != '{') { ps.skip(); continue; }               // non-dict object: skipped
                ++ps.p; ps.ws();
                KeySet okeys;
                int npts = -1;                                           // -1: no ptList seen (treated as [])
                double c[4] = {0, 0, 0, 0}; bool has[4] = {false, false, false, false}, isnull[4] = {false, false, false, false};
                bool pts_are_dicts = true;
                if (ps.p < ps.end && *ps.p == '}') { ++ps.p; continue; }
                do {
                    bool pl2; Span ok = ps.str(&pl2); ps.check_dup(okeys, ok); ps.colon();
                    if (pl2 && ps.key_is(ok, "polygon")) {
                        ps.need(ps.p < ps.end);
                        if (*ps.p != '{') Parser::fail();                // AttributeError in the reference: truncation -- slow lane decides
                        ++ps.p; ps.ws();
                        KeySet pkeys;
                        if (ps.p < ps.end && *ps.p == '}') { ++ps.p; continue; }
                        do {
                            bool pl3; Span pk = ps.str(&pl3); ps.check_dup(pkeys, pk); ps.colon();
                            if (pl3 && ps.key_is(pk, "ptList")) {
                                ps.need(ps.p < ps.end && *ps.p == '[');  // len() of other types: slow lane
                                ++ps.p; ps.ws();
                                npts = 0;
                                if (ps.p < ps.end && *ps.p == ']') { ++ps.p; continue; }
                                do {
                                    const int slot = npts++;
                                    ps.need(ps.p < ps.end);
                                    if (*ps.p != '{') { ps.skip(); pts_are_dicts = false; continue; }
                                    ++ps.p; ps.ws();
                                    KeySet qkeys;
                                    if (ps.p < ps.end && *ps.p == '}') { ++ps.p; continue; }
                                    do {
                                        bool pl4; Span qk = ps.str(&pl4); ps.check_dup(qkeys, qk); ps.colon();
                                        const bool isx = pl4 && ps.key_is(qk, "x"), isy = pl4 && ps.key_is(qk, "y");
                                        if ((isx || isy) && slot < 2) {
                                            ps.need(ps.p < ps.end);
                                            if (*ps.p == '{' || *ps.p == '[' || *ps.p == '"') Parser::fail();   // non-numeric coordinate: slow lane
                                            Span sp; double val; bool exact; Kind kd = ps.scalar(&sp, &val, &exact);
                                            // CPython computes IoU of int coordinates in exact big-int arithmetic; fp64 is the same only while
                                            // every product stays below 2^53, i.e. for |int| <= 2^25: larger ones take the CPython lane
                                            if (!exact || (kd == K_INT && std::fabs(val) > 33554432.0)) Parser::fail();
                                            const int idx = slot * 2 + (isy ? 1 : 0);
                                            has[idx] = true; c[idx] = val; isnull[idx] = kd == K_NULL;
                                        } else ps.skip();
                                    } while (ps.more('}'));
                                } while (ps.more(']'));
                            } else ps.skip();
                        } while (ps.more('}'));
                    } else ps.skip();
                } while (ps.more('}'));
                if (stop || npts != 2 || !pts_are_dicts || !(has[0] && has[1] && has[2] && has[3])) continue;
                if (isnull[0] || isnull[1] || isnull[2] || isnull[3]) {
                    // min(None, .) raises TypeError inside extract_boxes: the scan ends, the prefix is kept
                    out.boxes.insert(out.boxes.end(), {0.0, 0.0, 0.0, 0.0}); out.bvalid.push_back(0);
                    stop = true;
                    continue;
                }
                out.boxes.insert(out.boxes.end(), {c[0], c[1], c[2], c[3]}); out.bvalid.push_back(1);
            } while (ps.more(']'));
        } else ps.skip();
    } while (ps.more('}'));
    ps.ws();
    ps.need(ps.p == ps.end);
    out.status = ROW_OK;
}

// ---------------------------------------------------------------- step 5.5: object names
// replace_labels_by_mapping (processor.py:560-604) re-serialises every cell it can parse, so the native
// lane only takes cells that are already in json.dumps form (strict): the output is then the input with
// some "name" strings replaced.  Names that are not plain strings / None go to the Python lane.
static void parse_names_row(const char* text, size_t len, RowOut& out) {
    Parser ps{text, text, text + len, true};
    ps.need(ps.p < ps.end && *ps.p == '{');             // doc.get on a non-dict raises in the reference
    ++ps.p;
    bool have_list = false, not_a_list = false, first_member = true;
    KeySet top_keys;
    if (ps.p < ps.end && *ps.p == '}') { ++ps.p; }
    else do {
        const uint32_t key_at = ps.at();
        bool plain; Span k = ps.str(&plain); ps.check_dup(top_keys, k); ps.colon();
        const bool is_first = first_member;
        first_member = false;
        if (plain && ps.key_is(k, "objects")) {
            ps.need(ps.p < ps.end);
            if (*ps.p != '[') { ps.skip(); not_a_list = true; continue; }   // not a list: the cell is skipped by the reference
            have_list = true;
            ++ps.p;
            if (ps.p < ps.end && *ps.p == ']') { ++ps.p; }
            else do {
                ++out.list_len;
                ps.need(ps.p < ps.end);
                if (*ps.p != '{') { ps.skip(); continue; }               // non-dict elements are not counted
                const uint32_t obj_at = ps.at();
                ++ps.p;
                Span name{0, (uint32_t)-1};
                KeySet okeys;
                if (ps.p < ps.end && *ps.p == '}') { ++ps.p; }
                else do {
                    bool pl2; Span ok = ps.str(&pl2); ps.check_dup(okeys, ok); ps.colon();
                    if (pl2 && ps.key_is(ok, "name")) {
                        ps.need(ps.p < ps.end);
                        if (*ps.p == '"') { bool pl3; name = ps.str(&pl3); ps.need(pl3); }     // escapes: Python lane
                        else { Span sp; double v; ps.need(ps.scalar(&sp, &v) == K_NULL); }     // None; anything else: Python lane
                    } else ps.skip();
                } while (ps.more('}'));
                out.names.push_back(name);
                out.objspans.push_back(Span{obj_at, ps.at() - obj_at});
            } while (ps.more(']'));
            out.member = Span{key_at, ps.at() - key_at};                 // from the key's opening quote to the end of the value
            out.member_first = is_first;
            out.member_has_next = ps.p < ps.end && *ps.p == ',';
        } else ps.skip();
    } while (ps.more('}'));
    ps.need(ps.p == ps.end);
    out.status = have_list ? ROW_OK : (uint8_t)(not_a_list ? ROW_NOT_A_LIST : ROW_NO_LIST);
    if (!have_list) { out.names.clear(); out.objspans.clear(); out.list_len = 0; }
}

}  // namespace

struct dyd_ingest {
    int mode = 0;                            // 0 polygons (step 4), 1 boxes (step 5), 2 object names (step 5.5)
    int64_t n_rows = 0;
    std::vector<RowOut> rows;
    std::vector<int64_t> obj_base, vert_base;    // exclusive prefix over rows
    int64_t n_obj = 0, n_vert = 0, n_slow = 0, n_canon = 0;
};

namespace {

template <typename F>
void parallel_rows(int64_t n, int n_threads, F f) {
    if (n_threads <= 0) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
    n_threads = (int)std::min<int64_t>(n_threads, std::max<int64_t>(1, n / 256));
    if (n_threads <= 1) { f(0, n); return; }
    std::vector<std::thread> th;
    const int64_t per = (n + n_threads - 1) / n_threads;
    for (int t = 0; t < n_threads; ++t) {
        const int64_t a = t * per, b = std::min(n, a + per);
        if (a < b) th.emplace_back([=] { f(a, b); });
    }
    for (auto& t : th) t.join();
}

const char kNullPt[] = "[{\"x\": null, \"y\": null}, {\"x\": null, \"y\": null}]";

// bytes of the replacement ptList value for one polygon
inline size_t ptlist_value_len(const RowOut& r, const Poly& p, const int32_t* arg, const uint8_t* valid, int64_t q) {
    if (!valid[q] || p.v_count == 0) return sizeof(kNullPt) - 1;
    const Vertex* v = r.verts.data() + p.v_begin;
    return 32 + v[arg[4 * q]].sx.len + v[arg[4 * q + 1]].sy.len + v[arg[4 * q + 2]].sx.len + v[arg[4 * q + 3]].sy.len;
}
inline char* put(char* o, const char* s, size_t n) { memcpy(o, s, n); return o + n; }
inline char* ptlist_value_write(char* o, const char* text, const RowOut& r, const Poly& p, const int32_t* arg,
                                const uint8_t* valid, int64_t q) {
    if (!valid[q] || p.v_count == 0) return put(o, kNullPt, sizeof(kNullPt) - 1);
    const Vertex* v = r.verts.data() + p.v_begin;
    const Span a = v[arg[4 * q]].sx, b = v[arg[4 * q + 1]].sy, c = v[arg[4 * q + 2]].sx, d = v[arg[4 * q + 3]].sy;
    o = put(o, "[{\"x\": ", 7); o = put(o, text + a.off, a.len);
    o = put(o, ", \"y\": ", 7); o = put(o, text + b.off, b.len);
    o = put(o, "}, {\"x\": ", 9); o = put(o, text + c.off, c.len);
    o = put(o, ", \"y\": ", 7); o = put(o, text + d.off, d.len);
    return put(o, "}]", 2);
}
inline size_t wrap_len(const Poly& p) {      // extra bytes around the value for appended keys
    if (p.splice_kind == 0) return 0;
    const size_t lead = p.container_empty ? 0 : 2;                       // ", "
    return lead + (p.splice_kind == 1 ? 10 : 23);                        // "ptList":_  |  "polygon": {"ptList":_ ... }
}

}  // namespace

extern "C" int dyd_ingest_cells(const uint8_t* text, const int64_t* off, const uint8_t* is_text, int64_t n_rows,
                                int mode, int n_threads, dyd_ingest** out) {
    if (!out || n_rows < 0 || (n_rows > 0 && (!text || !off)) || (mode < 0 || mode > 2)) return DYD_E_ARG;
    dyd_ingest* h = new dyd_ingest();
    h->mode = mode; h->n_rows = n_rows;
    h->rows.resize((size_t)n_rows);
    parallel_rows(n_rows, n_threads, [&](int64_t a, int64_t b) {
        for (int64_t r = a; r < b; ++r) {
            RowOut& ro = h->rows[(size_t)r];
            if (is_text && !is_text[r]) { ro.status = ROW_NOT_TEXT; continue; }
            const char* s = reinterpret_cast<const char*>(text) + off[r];
            const size_t len = (size_t)(off[r + 1] - off[r]);
            try {
                if (len >= (1ull << 31)) throw Fail{};
                if (mode == 2 && len == 0) { ro.status = ROW_NOT_TEXT; continue; }     // an empty cell is skipped like a missing one
                if (mode == 1) { parse_box_row(s, len, ro); continue; }
                try {
                    if (mode == 0) parse_polygon_row(s, len, true, ro); else parse_names_row(s, len, ro);
                } catch (const Fail&) {
                    // not in json.dumps form (or not acceptable at all): rewrite it canonically and try again
                    ro = RowOut();
                    std::string canon;
                    canonicalize(s, len, canon);
                    if (canon.size() >= (1ull << 31)) throw Fail{};
                    if (mode == 0) parse_polygon_row(canon.data(), canon.size(), true, ro); else parse_names_row(canon.data(), canon.size(), ro);
                    ro.canon = std::move(canon);
                }
            } catch (const Fail&) {
                ro = RowOut(); ro.status = ROW_SLOW;
            } catch (const std::bad_alloc&) {
                ro = RowOut(); ro.status = ROW_SLOW;
            }
        }
    });
    h->obj_base.resize((size_t)n_rows + 1); h->vert_base.resize((size_t)n_rows + 1);
    int64_t no = 0, nv = 0, ns = 0;
    for (int64_t r = 0; r < n_rows; ++r) {
        const RowOut& ro = h->rows[(size_t)r];
        h->obj_base[(size_t)r] = no; h->vert_base[(size_t)r] = nv;
        if (ro.status == ROW_SLOW) ++ns;
        if (!ro.canon.empty()) ++h->n_canon;
        no += mode == 0 ? (int64_t)ro.polys.size() : (mode == 1 ? (int64_t)ro.bvalid.size() : (int64_t)ro.names.size());
        nv += (int64_t)ro.verts.size();
    }
    h->obj_base[(size_t)n_rows] = no; h->vert_base[(size_t)n_rows] = nv;
    h->n_obj = no; h->n_vert = nv; h->n_slow = ns;
    *out = h;
    return 0;
}

extern "C" void dyd_ingest_free(dyd_ingest* h) { delete h; }

// Rows that were valid JSON in another style have been rewritten canonically; their spans refer to the
// rewritten text.  n_canon = how many; with new_text == NULL new_off[n+1] receives the offsets of the
// effective texts (rewritten where rewritten, input otherwise), with new_text they are written.  The caller
// passes these buffers, not the input, to every export / egress call that takes `text` and `off`.
extern "C" int dyd_ingest_effective_text(const dyd_ingest* h, const uint8_t* text, const int64_t* off, int64_t* n_canon,
                                         int64_t* new_off, uint8_t* new_text, int n_threads) {
    if (!h) return DYD_E_ARG;
    if (n_canon) *n_canon = h->n_canon;
    if (!new_off) return 0;
    if (!off || (!text && h->n_rows > 0)) return DYD_E_ARG;
    const int64_t n = h->n_rows;
    if (!new_text) {
        new_off[0] = 0;
        for (int64_t r = 0; r < n; ++r) {
            const RowOut& ro = h->rows[(size_t)r];
            new_off[r + 1] = new_off[r] + (ro.canon.empty() ? off[r + 1] - off[r] : (int64_t)ro.canon.size());
        }
        return 0;
    }
    parallel_rows(n, n_threads, [&](int64_t a, int64_t b) {
        for (int64_t r = a; r < b; ++r) {
            const RowOut& ro = h->rows[(size_t)r];
            if (ro.canon.empty()) memcpy(new_text + new_off[r], text + off[r], (size_t)(off[r + 1] - off[r]));
            else memcpy(new_text + new_off[r], ro.canon.data(), ro.canon.size());
        }
    });
    return 0;
}

// json.dumps(json.loads(text), ensure_ascii=False) of one document (test hook of the canonical rewriter).
// Returns the length, -1 if this file leaves the document to CPython, -2 if `cap` is too small.
extern "C" int64_t dyd_json_canonical(const uint8_t* text, int64_t len, uint8_t* out, int64_t cap) {
    if (!text || len < 0) return -1;
    try {
        std::string o;
        canonicalize(reinterpret_cast<const char*>(text), (size_t)len, o);
        if ((int64_t)o.size() > cap || !out) return -2;
        memcpy(out, o.data(), o.size());
        return (int64_t)o.size();
    } catch (const Fail&) {
        return -1;
    } catch (const std::bad_alloc&) {
        return -1;
    }
}

extern "C" int dyd_ingest_sizes(const dyd_ingest* h, int64_t* n_obj, int64_t* n_vert, int64_t* n_slow) {
    if (!h) return DYD_E_ARG;
    if (n_obj) *n_obj = h->n_obj;
    if (n_vert) *n_vert = h->n_vert;
    if (n_slow) *n_slow = h->n_slow;
    return 0;
}

// mode 0: status[n], img_off[n+1], poly_off[n_obj+1], xy[2*n_vert], wh_off[2n] (-1 absent) / wh_len[2n] / wh_kind[2n]
extern "C" int dyd_ingest_export_polygons(const dyd_ingest* h, uint8_t* status, int64_t* img_off, int64_t* poly_off, double* xy,
                                          int64_t* wh_off, int32_t* wh_len, uint8_t* wh_kind, int n_threads) {
    if (!h || h->mode != 0 || !status || !img_off || !poly_off) return DYD_E_ARG;
    const int64_t n = h->n_rows;
    img_off[n] = h->n_obj; poly_off[h->n_obj] = h->n_vert;
    parallel_rows(n, n_threads, [&](int64_t a, int64_t b) {
        for (int64_t r = a; r < b; ++r) {
            const RowOut& ro = h->rows[(size_t)r];
            status[r] = ro.status;
            img_off[r] = h->obj_base[(size_t)r];
            int64_t q = h->obj_base[(size_t)r]; const int64_t vb = h->vert_base[(size_t)r];
            for (const Poly& p : ro.polys) poly_off[q++] = vb + p.v_begin;
            if (xy) { double* o = xy + 2 * vb; for (const Vertex& v : ro.verts) { *o++ = v.x; *o++ = v.y; } }
            if (wh_off) {
                wh_off[2 * r] = ro.width.len ? (int64_t)ro.width.off : -1; wh_len[2 * r] = (int32_t)ro.width.len; wh_kind[2 * r] = ro.width_kind;
                wh_off[2 * r + 1] = ro.height.len ? (int64_t)ro.height.off : -1; wh_len[2 * r + 1] = (int32_t)ro.height.len; wh_kind[2 * r + 1] = ro.height_kind;
            }
        }
    });
    return 0;
}

// mode 1: status[n], img_off[n+1], pts[4*n_obj], valid[n_obj]
extern "C" int dyd_ingest_export_boxes(const dyd_ingest* h, uint8_t* status, int64_t* img_off, double* pts, uint8_t* valid, int n_threads) {
    if (!h || h->mode != 1 || !status || !img_off) return DYD_E_ARG;
    const int64_t n = h->n_rows;
    img_off[n] = h->n_obj;
    parallel_rows(n, n_threads, [&](int64_t a, int64_t b) {
        for (int64_t r = a; r < b; ++r) {
            const RowOut& ro = h->rows[(size_t)r];
            status[r] = ro.status;
            const int64_t q = h->obj_base[(size_t)r];
            img_off[r] = q;
            if (!ro.bvalid.empty()) {
                memcpy(pts + 4 * q, ro.boxes.data(), sizeof(double) * ro.boxes.size());
                memcpy(valid + q, ro.bvalid.data(), ro.bvalid.size());
            }
        }
    });
    return 0;
}

// Output cell texts of step 4 for the ROW_OK rows (other rows get length 0).  Two calls: with out == NULL the
// per-row lengths are written to out_off[1..n] as an inclusive prefix (out_off[0] = 0); then with out != NULL.
extern "C" int dyd_egress_ptlist(const dyd_ingest* h, const uint8_t* text, const int64_t* off, const int32_t* arg,
                                 const uint8_t* valid, int64_t* out_off, uint8_t* out, int n_threads) {
    if (!h || h->mode != 0 || !text || !off || !out_off || (h->n_obj > 0 && (!arg || !valid))) return DYD_E_ARG;
    const int64_t n = h->n_rows;
    if (!out) {
        parallel_rows(n, n_threads, [&](int64_t a, int64_t b) {
            for (int64_t r = a; r < b; ++r) {
                const RowOut& ro = h->rows[(size_t)r];
                int64_t len = 0;
                if (ro.status == ROW_OK) {
                    len = off[r + 1] - off[r];
                    int64_t q = h->obj_base[(size_t)r];
                    for (const Poly& p : ro.polys) { len += (int64_t)ptlist_value_len(ro, p, arg, valid, q) + (int64_t)wrap_len(p) - (int64_t)p.splice_len; ++q; }
                }
                out_off[r + 1] = len;
            }
        });
        out_off[0] = 0;
        for (int64_t r = 0; r < n; ++r) out_off[r + 1] += out_off[r];
        return 0;
    }
    parallel_rows(n, n_threads, [&](int64_t a, int64_t b) {
        for (int64_t r = a; r < b; ++r) {
            const RowOut& ro = h->rows[(size_t)r];
            if (ro.status != ROW_OK) continue;
            const char* src = reinterpret_cast<const char*>(text) + off[r];
            const uint32_t len = (uint32_t)(off[r + 1] - off[r]);
            char* o = reinterpret_cast<char*>(out) + out_off[r];
            uint32_t cur = 0;
            int64_t q = h->obj_base[(size_t)r];
            for (const Poly& p : ro.polys) {
                o = put(o, src + cur, p.splice_off - cur);
                if (p.splice_kind == 0) {
                    o = ptlist_value_write(o, src, ro, p, arg, valid, q);
                } else {
                    if (!p.container_empty) o = put(o, ", ", 2);
                    if (p.splice_kind == 2) o = put(o, "\"polygon\": {", 12);
                    o = put(o, "\"ptList\": ", 10);
                    o = ptlist_value_write(o, src, ro, p, arg, valid, q);
                    if (p.splice_kind == 2) o = put(o, "}", 1);
                }
                cur = p.splice_off + p.splice_len;
                ++q;
            }
            o = put(o, src + cur, len - cur);
        }
    });
    return 0;
}

extern "C" int dyd_py_float_repr(double v, char* out40) { return py_float_repr(v, out40); }

// ---------------------------------------------------------------- CSV egress
// Body of DataFrame.to_csv(index=False) for frames made of string / float64 / int64 / bool columns,
// byte-identical to pandas (which drives CPython's csv.writer with QUOTE_MINIMAL): a field is quoted
// iff it contains ',', '"', '\n' or '\r' (quotes doubled); missing values are empty; floats are
// repr(float); rows end with '\n'.  The header line (and BOM) is written by the caller.
// Column kinds: 0 string (Arrow large_string + validity bytes), 1 float64, 2 int64, 3 bool (uint8).
namespace {

struct CsvCol { int kind; const int64_t* off; const uint8_t* data; const uint8_t* valid; };

inline size_t csv_field_len(const CsvCol& c, int64_t r) {
    switch (c.kind) {
        case 0: {
            if (c.valid && !c.valid[r]) return 0;
            const uint8_t* s = c.data + c.off[r]; const int64_t n = c.off[r + 1] - c.off[r];
            size_t q = 0; bool need = false;
            dyd_simd::csv_scan(s, n, q, need);
            return (size_t)n + (need ? q + 2 : 0);
        }
        case 1: {
            const double v = reinterpret_cast<const double*>(c.data)[r];
            if (v != v) return 0;
            if (std::isinf(v)) return v > 0 ? 3 : 4;                 // str(float): "inf" / "-inf" (JSON spells it Infinity)
            char b[40]; return (size_t)py_float_repr(v, b);
        }
        case 2: { char b[24]; auto e = std::to_chars(b, b + 24, reinterpret_cast<const int64_t*>(c.data)[r]); return (size_t)(e.ptr - b); }
        default: return c.data[r] ? 4 : 5;
    }
}
// `slack`: 64 writable bytes follow whatever this call produces (lets the quote-doubling copy use wide stores)
inline char* csv_field_write(char* o, const CsvCol& c, int64_t r, bool slack = false) {
    switch (c.kind) {
        case 0: {
            if (c.valid && !c.valid[r]) return o;
            const uint8_t* s = c.data + c.off[r]; const int64_t n = c.off[r + 1] - c.off[r];
            size_t q = 0; bool need = false;
            dyd_simd::csv_scan(s, n, q, need);
            if (!need) { memcpy(o, s, (size_t)n); return o + n; }
            *o++ = '"';
            o = q ? dyd_simd::csv_double_quotes(o, s, n, slack) : (char*)memcpy(o, s, (size_t)n) + n;
            *o++ = '"';
            return o;
        }
        case 1: {
            const double v = reinterpret_cast<const double*>(c.data)[r];
            if (v != v) return o;
            if (std::isinf(v)) { if (v > 0) { memcpy(o, "inf", 3); return o + 3; } memcpy(o, "-inf", 4); return o + 4; }
            return o + py_float_repr(v, o);
        }
        case 2: { auto e = std::to_chars(o, o + 24, reinterpret_cast<const int64_t*>(c.data)[r]); return e.ptr; }
        default: if (c.data[r]) { memcpy(o, "True", 4); return o + 4; } memcpy(o, "False", 5); return o + 5;
    }
}

}  // namespace

// kinds[c], offs[c] (string columns), datas[c], valids[c] (may be NULL).  With out == NULL fills row_off
// (int64[n_rows+1], exclusive prefix of row byte lengths); with out != NULL writes the rows.
extern "C" int dyd_csv_write(const int32_t* kinds, const int64_t* const* offs, const uint8_t* const* datas,
                             const uint8_t* const* valids, int32_t n_cols, int64_t n_rows, int64_t* row_off,
                             uint8_t* out, int n_threads) {
    if (!kinds || !datas || !row_off || n_cols < 2 || n_rows < 0) return DYD_E_ARG;     // single-column frames quote empty fields
    std::vector<CsvCol> cols((size_t)n_cols);
    for (int c = 0; c < n_cols; ++c) {
        cols[(size_t)c] = CsvCol{kinds[c], offs ? offs[c] : nullptr, datas[c], valids ? valids[c] : nullptr};
        if (kinds[c] < 0 || kinds[c] > 3 || !datas[c] || (kinds[c] == 0 && !cols[(size_t)c].off)) return DYD_E_ARG;
    }
    if (!out) {
        parallel_rows(n_rows, n_threads, [&](int64_t a, int64_t b) {
            for (int64_t r = a; r < b; ++r) {
                size_t len = (size_t)n_cols;                      // commas + newline
                for (const CsvCol& c : cols) len += csv_field_len(c, r);
                row_off[r + 1] = (int64_t)len;
            }
        });
        row_off[0] = 0;
        for (int64_t r = 0; r < n_rows; ++r) row_off[r + 1] += row_off[r];
        return 0;
    }
    parallel_rows(n_rows, n_threads, [&](int64_t a, int64_t b) {
        for (int64_t r = a; r < b; ++r) {
            char* o = reinterpret_cast<char*>(out) + row_off[r];
            for (int c = 0; c < n_cols; ++c) { if (c) *o++ = ','; o = csv_field_write(o, cols[(size_t)c], r); }
            *o++ = '\n';
        }
    });
    return 0;
}

// ---- to_csv straight into the file --------------------------------------------------------------------
// DataFrame.to_csv(path, index=False) of selected rows without materialising either the selected frame or the
// whole body: worker threads format blocks of rows into their own buffers (quote doubling with wide stores),
// the calling thread writes finished blocks in order with write(2) -- formatting and the page-cache copy overlap,
// and a buffered write to one file does not scale over threads anyway (the inode lock serialises it).
namespace {

struct OutBuf {
    char* p = nullptr; size_t cap = 0, len = 0;
    ~OutBuf() { free(p); }
    inline char* room(size_t need) {                  // `need` bytes + 64 of slack behind them
        if (cap - len < need + 64) {
            size_t nc = std::max(cap * 2, len + need + 64 + (1u << 16));
            char* np = (char*)realloc(p, nc);
            if (!np) throw std::bad_alloc();
            p = np; cap = nc;
        }
        return p + len;
    }
    void release() { free(p); p = nullptr; cap = len = 0; }
};

inline void csv_format_rows(const std::vector<CsvCol>& cols, const int64_t* rows, int64_t a, int64_t b, OutBuf& out) {
    const int nc = (int)cols.size();
    for (int64_t k = a; k < b; ++k) {
        const int64_t r = rows ? rows[k] : k;
        for (int c = 0; c < nc; ++c) {
            const CsvCol& col = cols[(size_t)c];
            size_t worst = 48;                                                   // number / bool + separator
            if (col.kind == 0) worst = 2 * (size_t)(col.off[r + 1] - col.off[r]) + 4;
            char* o = out.room(worst);
            char* const o0 = o;
            if (c) *o++ = ',';
            o = csv_field_write(o, col, r, true);
            out.len += (size_t)(o - o0);
        }
        *out.room(1) = '\n'; ++out.len;
    }
}

inline bool write_all(int fd, const char* p, size_t n) {
    while (n) {
        const ssize_t w = ::write(fd, p, n);
        if (w < 0) { if (errno == EINTR) continue; return false; }
        p += w; n -= (size_t)w;
    }
    return true;
}

}  // namespace

extern "C" int dyd_csv_write_file(const char* path, int32_t append, const uint8_t* prefix, int64_t prefix_len,
                                  const int32_t* kinds, const int64_t* const* offs, const uint8_t* const* datas,
                                  const uint8_t* const* valids, int32_t n_cols, const int64_t* rows, int64_t n_sel,
                                  int n_threads, int64_t* bytes_written) {
    if (!path || !kinds || !datas || n_cols < 2 || n_sel < 0 || prefix_len < 0 || (prefix_len > 0 && !prefix)) return DYD_E_ARG;
    std::vector<CsvCol> cols((size_t)n_cols);
    double str_bytes = 0;
    for (int c = 0; c < n_cols; ++c) {
        cols[(size_t)c] = CsvCol{kinds[c], offs ? offs[c] : nullptr, datas[c], valids ? valids[c] : nullptr};
        if (kinds[c] < 0 || kinds[c] > 3 || !datas[c] || (kinds[c] == 0 && !cols[(size_t)c].off)) return DYD_E_ARG;
    }
    // rows per block: about 2 MB of output, estimated from a sample of the selected rows
    {
        const int64_t step = std::max<int64_t>(1, n_sel / 64);
        int64_t cnt = 0;
        for (int64_t k = 0; k < n_sel; k += step, ++cnt) {
            const int64_t r = rows ? rows[k] : k;
            for (const CsvCol& c : cols) str_bytes += c.kind == 0 ? (double)(c.off[r + 1] - c.off[r]) + 1 : 12;
        }
        if (cnt) str_bytes /= (double)cnt;
    }
    const int64_t block_rows = std::max<int64_t>(16, (int64_t)((2 << 20) / std::max(str_bytes, 1.0)));
    const int64_t nb = (n_sel + block_rows - 1) / block_rows;
    // An existing file is overwritten in place and cut to the new length at the end: the page-cache pages of the old
    // content are reused instead of being freed and allocated again (1.6x faster for a re-run of a step; the final
    // state is that of open(path, "wb") + write).
    const int fd = ::open(path, O_WRONLY | O_CREAT | O_CLOEXEC | (append ? O_APPEND : 0), 0666);
    if (fd < 0) return DYD_E_IO;
    int64_t total = 0;
    bool ok = prefix_len == 0 || write_all(fd, (const char*)prefix, (size_t)prefix_len);
    total += prefix_len;
    int T = n_threads > 0 ? n_threads : (int)std::max(1u, std::thread::hardware_concurrency());
    T = (int)std::min<int64_t>(T, nb);
    try {
        if (ok && nb > 0 && T <= 1) {
            OutBuf buf;
            for (int64_t b = 0; b < nb && ok; ++b) {
                buf.len = 0;
                csv_format_rows(cols, rows, b * block_rows, std::min(n_sel, (b + 1) * block_rows), buf);
                ok = write_all(fd, buf.p, buf.len);
                total += (int64_t)buf.len;
            }
        } else if (ok && nb > 0) {
            std::vector<OutBuf> slot((size_t)nb);
            std::vector<uint8_t> ready((size_t)nb, 0);
            std::mutex mu;
            std::condition_variable cv;
            std::atomic<int64_t> next{0};
            int64_t written = 0;
            bool failed = false;
            const int64_t window = 2 * (int64_t)T + 2;                  // blocks in flight: bounds the memory held
            std::vector<std::thread> th;
            for (int t = 0; t < T; ++t)
                th.emplace_back([&] {
                    for (;;) {
                        const int64_t b = next.fetch_add(1);
                        if (b >= nb) return;
                        {
                            std::unique_lock<std::mutex> lk(mu);
                            cv.wait(lk, [&] { return failed || b < written + window; });
                            if (failed) return;
                        }
                        bool good = true;
                        try { csv_format_rows(cols, rows, b * block_rows, std::min(n_sel, (b + 1) * block_rows), slot[(size_t)b]); }
                        catch (const std::bad_alloc&) { good = false; }
                        {
                            std::lock_guard<std::mutex> lk(mu);
                            if (!good) failed = true;
                            ready[(size_t)b] = 1;
                        }
                        cv.notify_all();
                    }
                });
            for (int64_t b = 0; b < nb; ++b) {
                {
                    std::unique_lock<std::mutex> lk(mu);
                    cv.wait(lk, [&] { return failed || ready[(size_t)b]; });
                    if (failed) break;
                }
                const bool w = write_all(fd, slot[(size_t)b].p, slot[(size_t)b].len);
                total += (int64_t)slot[(size_t)b].len;
                slot[(size_t)b].release();
                {
                    std::lock_guard<std::mutex> lk(mu);
                    if (!w) failed = true;
                    written = b + 1;
                }
                cv.notify_all();
                if (!w) break;
            }
            for (auto& t : th) t.join();
            ok = !failed;
        }
    } catch (const std::bad_alloc&) {
        ok = false;
    }
    if (ok && !append && ::ftruncate(fd, (off_t)total) != 0) ok = false;
    const int saved = errno;
    if (::close(fd) != 0) ok = false;
    if (!ok) { errno = saved; return DYD_E_IO; }
    if (bytes_written) *bytes_written = total;
    return 0;
}

// mode 2: status[n], cell_off[n+1] (objects per cell), name_off[n_obj] (byte offset inside the cell's text) /
// name_len[n_obj] (-1: the object has no name or None)
extern "C" int dyd_ingest_export_names(const dyd_ingest* h, uint8_t* status, int64_t* cell_off, int64_t* name_off, int32_t* name_len,
                                       int n_threads) {
    if (!h || h->mode != 2 || !status || !cell_off) return DYD_E_ARG;
    const int64_t n = h->n_rows;
    cell_off[n] = h->n_obj;
    parallel_rows(n, n_threads, [&](int64_t a, int64_t b) {
        for (int64_t r = a; r < b; ++r) {
            const RowOut& ro = h->rows[(size_t)r];
            status[r] = ro.status;
            int64_t q = h->obj_base[(size_t)r];
            cell_off[r] = q;
            for (const Span& sp : ro.names) { if (name_off) name_off[q] = sp.off; if (name_len) name_len[q] = (int32_t)sp.len; ++q; }
        }
    });
    return 0;
}

// New cell texts of step 5.5: the input text with the name of every object whose flag is set replaced by
// vocab entry obj_new[q] (bytes already JSON-escaped by the caller, without the quotes).  Cells that are
// not ROW_OK get length 0.  out == NULL: fill out_off[n+1]; else write.
extern "C" int dyd_egress_names(const dyd_ingest* h, const uint8_t* text, const int64_t* off, const uint8_t* obj_flag,
                                const int32_t* obj_new, const uint8_t* vocab_bytes, const int64_t* vocab_off, int64_t n_vocab,
                                int64_t* out_off, uint8_t* out, int n_threads) {
    if (!h || h->mode != 2 || !text || !off || !out_off || (h->n_obj > 0 && (!obj_flag || !obj_new || !vocab_bytes || !vocab_off))) return DYD_E_ARG;
    const int64_t n = h->n_rows;
    auto cell = [&](int64_t r, uint8_t* dst) -> int64_t {
        const RowOut& ro = h->rows[(size_t)r];
        if (ro.status != ROW_OK) return 0;
        const uint8_t* src = text + off[r];
        const int64_t len = off[r + 1] - off[r];
        int64_t q = h->obj_base[(size_t)r], pos = 0, total = 0;
        for (const Span& sp : ro.names) {
            if (obj_flag[q] && sp.len != (uint32_t)-1) {
                const int32_t v = obj_new[q];
                if (v < 0 || v >= n_vocab) return -1;
                const int64_t nl = vocab_off[v + 1] - vocab_off[v];
                if (dst) { memcpy(dst + total, src + pos, (size_t)(sp.off - pos)); memcpy(dst + total + (sp.off - pos), vocab_bytes + vocab_off[v], (size_t)nl); }
                total += (sp.off - pos) + nl;
                pos = (int64_t)sp.off + sp.len;
            }
            ++q;
        }
        if (dst) memcpy(dst + total, src + pos, (size_t)(len - pos));
        return total + (len - pos);
    };
    bool bad = false;
    if (!out) {
        out_off[0] = 0;
        parallel_rows(n, n_threads, [&](int64_t a, int64_t b) { for (int64_t r = a; r < b; ++r) { const int64_t l = cell(r, nullptr); if (l < 0) bad = true; out_off[r + 1] = l < 0 ? 0 : l; } });
        for (int64_t r = 0; r < n; ++r) out_off[r + 1] += out_off[r];
        return bad ? DYD_E_ARG : 0;
    }
    parallel_rows(n, n_threads, [&](int64_t a, int64_t b) { for (int64_t r = a; r < b; ++r) cell(r, out + out_off[r]); });
    return 0;
}

// mode 2, for the split (step 6): list_len[n] (elements of "objects", dicts or not), obj_off / obj_len [n_obj]
// (span of every dict object inside its cell's text)
extern "C" int dyd_ingest_export_objects(const dyd_ingest* h, int32_t* list_len, int64_t* obj_off, int32_t* obj_len, int n_threads) {
    if (!h || h->mode != 2 || !list_len) return DYD_E_ARG;
    parallel_rows(h->n_rows, n_threads, [&](int64_t a, int64_t b) {
        for (int64_t r = a; r < b; ++r) {
            const RowOut& ro = h->rows[(size_t)r];
            list_len[r] = ro.list_len;
            int64_t q = h->obj_base[(size_t)r];
            for (const Span& sp : ro.objspans) { if (obj_off) obj_off[q] = sp.off; if (obj_len) obj_len[q] = (int32_t)sp.len; ++q; }
        }
    });
    return 0;
}

// Cells of the split's expanded rows (processor.py:760-775): json.dumps of the document with "objects" replaced
// by the single object exp_obj[i] (global dict-object index) renamed to label exp_tok[i] -- in text form: the
// document without its "objects" member, then `"objects": [` + the object with its name spliced + `]}`.
// Labels are passed JSON-escaped.  out == NULL: fill out_off[n_exp+1]; else write.
extern "C" int dyd_egress_split(const dyd_ingest* h, const uint8_t* text, const int64_t* off, int64_t n_exp,
                                const int64_t* exp_cell, const int64_t* exp_obj, const int32_t* exp_tok,
                                const uint8_t* tok_bytes, const int64_t* tok_off, int64_t n_tok,
                                int64_t* out_off, uint8_t* out, int n_threads) {
    if (!h || h->mode != 2 || n_exp < 0 || !out_off || (n_exp > 0 && (!text || !off || !exp_cell || !exp_obj || !exp_tok || !tok_bytes || !tok_off)))
        return DYD_E_ARG;
    static const char kOpen[] = "\"objects\": [";
    bool bad = false;
    auto one = [&](int64_t i, uint8_t* dst) -> int64_t {
        const int64_t r = exp_cell[i];
        if (r < 0 || r >= h->n_rows) return -1;
        const RowOut& ro = h->rows[(size_t)r];
        const int64_t k = exp_obj[i] - h->obj_base[(size_t)r];
        const int32_t t = exp_tok[i];
        if (ro.status != ROW_OK || k < 0 || k >= (int64_t)ro.objspans.size() || t < 0 || t >= n_tok) return -1;
        const Span os = ro.objspans[(size_t)k], ns = ro.names[(size_t)k];
        if (ns.len == (uint32_t)-1) return -1;                       // an object without a name has no labels
        const uint8_t* src = text + off[r];
        const int64_t len = off[r + 1] - off[r];
        // the document without the "objects" member and without its closing brace
        int64_t cut_a = ro.member.off, cut_b = (int64_t)ro.member.off + ro.member.len;
        if (!ro.member_first) cut_a -= 2; else if (ro.member_has_next) cut_b += 2;
        const bool others = !(ro.member_first && !ro.member_has_next);
        const int64_t tl = tok_off[t + 1] - tok_off[t];
        uint8_t* o = dst;
        int64_t total = 0;
        auto put = [&](const void* p, int64_t n) { if (o) { memcpy(o + total, p, (size_t)n); } total += n; };
        put(src, cut_a);
        put(src + cut_b, (len - 1) - cut_b);
        if (others) put(", ", 2);
        put(kOpen, (int64_t)sizeof(kOpen) - 1);
        put(src + os.off, (int64_t)ns.off - os.off);
        put(tok_bytes + tok_off[t], tl);
        put(src + ns.off + ns.len, ((int64_t)os.off + os.len) - ((int64_t)ns.off + ns.len));
        put("]}", 2);
        return total;
    };
    if (!out) {
        out_off[0] = 0;
        parallel_rows(n_exp, n_threads, [&](int64_t a, int64_t b) { for (int64_t i = a; i < b; ++i) { const int64_t l = one(i, nullptr); if (l < 0) bad = true; out_off[i + 1] = l < 0 ? 0 : l; } });
        for (int64_t i = 0; i < n_exp; ++i) out_off[i + 1] += out_off[i];
        return bad ? DYD_E_ARG : 0;
    }
    parallel_rows(n_exp, n_threads, [&](int64_t a, int64_t b) { for (int64_t i = a; i < b; ++i) one(i, out + out_off[i]); });
    return 0;
}

// ------------------------------------------------------------------------------------------------
// YOLO label text (processor.py:1045-1052): per image, one line per kept box
//     f"{cls} {cx:.6f} {cy:.6f} {bw:.6f} {bh:.6f}"      joined by "\n" (no trailing newline)
// from the device-computed cx/cy/w/h (dyd_yolo_normalise).  "%.6f" of glibc and of CPython are both
// the correctly rounded decimal expansion; nan / inf are spelled the CPython way.
// ------------------------------------------------------------------------------------------------
namespace {
inline char* put_f6(char* o, double v) {
    if (std::isnan(v)) { memcpy(o, "nan", 3); return o + 3; }
    if (std::isinf(v)) { if (v < 0) *o++ = '-'; memcpy(o, "inf", 3); return o + 3; }
    return o + snprintf(o, 330, "%.6f", v);
}
}  // namespace

extern "C" int dyd_yolo_format(const int64_t* img_off, const int32_t* class_id, const double* cxcywh, const uint8_t* ok,
                               int64_t n_img, int64_t* out_off, uint8_t* out, int n_threads) {
    if (n_img < 0 || (n_img > 0 && (!img_off || !out_off)) ) return DYD_E_ARG;
    if (n_img == 0) { if (out_off) out_off[0] = 0; return 0; }
    if (img_off[n_img] > img_off[0] && (!class_id || !cxcywh || !ok)) return DYD_E_ARG;
    auto image = [&](int64_t i, char* dst) -> int64_t {          // writes (dst != NULL) or measures one image
        char line[4 * 340 + 32];
        int64_t total = 0;
        bool first = true;
        for (int64_t q = img_off[i]; q < img_off[i + 1]; ++q) {
            if (!ok[q]) continue;
            char* o = line;
            if (!first) *o++ = '\n';
            first = false;
            o += snprintf(o, 16, "%d", (int)class_id[q]);
            for (int k = 0; k < 4; ++k) { *o++ = ' '; o = put_f6(o, cxcywh[4 * q + k]); }
            const int64_t len = o - line;
            if (dst) memcpy(dst + total, line, (size_t)len);
            total += len;
        }
        return total;
    };
    if (!out) {
        out_off[0] = 0;
        parallel_rows(n_img, n_threads, [&](int64_t a, int64_t b) { for (int64_t i = a; i < b; ++i) out_off[i + 1] = image(i, nullptr); });
        for (int64_t i = 0; i < n_img; ++i) out_off[i + 1] += out_off[i];
        return 0;
    }
    parallel_rows(n_img, n_threads, [&](int64_t a, int64_t b) { for (int64_t i = a; i < b; ++i) image(i, (char*)out + out_off[i]); });
    return 0;
}

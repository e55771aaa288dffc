// Byte-scanning helpers of the host lanes (CSV writer / reader), AVX-512 where the host CPU has it.
//
// Every helper has a portable scalar form with the same result; the wide forms are picked at run time
// (__builtin_cpu_supports), never at compile time, because libdyd.so is built on one machine and runs on another.
// DYD_NO_SIMD=1 forces the scalar forms (the tests run both and compare).
#pragma once
#include <immintrin.h>

#include <cstdint>
#include <cstdlib>
#include <cstring>

namespace dyd_simd {

inline bool wide() {
    static const bool v = [] {
        const char* e = getenv("DYD_NO_SIMD");
        if (e && *e && *e != '0') return false;
        __builtin_cpu_init();
        return __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vl") &&
               __builtin_cpu_supports("avx512vbmi2") && __builtin_cpu_supports("bmi2") && __builtin_cpu_supports("popcnt");
    }();
    return v;
}

#define DYD_AVX512 __attribute__((target("avx512f,avx512bw,avx512vl,avx512vbmi2,bmi2,popcnt")))

// ---- CSV QUOTE_MINIMAL: does the field need quotes, and how many '"' does it hold -------------------------
inline void csv_scan_scalar(const uint8_t* s, int64_t n, size_t& quotes, bool& need) {
    size_t q = 0; bool nd = false;
    for (int64_t i = 0; i < n; ++i) { const uint8_t ch = s[i]; if (ch == '"') { ++q; nd = true; } else if (ch == ',' || ch == '\n' || ch == '\r') nd = true; }
    quotes = q; need = nd;
}
DYD_AVX512 inline void csv_scan_wide(const uint8_t* s, int64_t n, size_t& quotes, bool& need) {
    const __m512i vq = _mm512_set1_epi8('"'), vc = _mm512_set1_epi8(','), vn = _mm512_set1_epi8('\n'), vr = _mm512_set1_epi8('\r');
    size_t q = 0; uint64_t other = 0;
    int64_t i = 0;
    for (; i + 64 <= n; i += 64) {
        const __m512i v = _mm512_loadu_si512(s + i);
        q += (size_t)_mm_popcnt_u64(_mm512_cmpeq_epi8_mask(v, vq));
        other |= _mm512_cmpeq_epi8_mask(v, vc) | _mm512_cmpeq_epi8_mask(v, vn) | _mm512_cmpeq_epi8_mask(v, vr);
    }
    if (i < n) {
        const __mmask64 m = (~0ULL) >> (64 - (n - i));
        const __m512i v = _mm512_maskz_loadu_epi8(m, s + i);
        q += (size_t)_mm_popcnt_u64(_mm512_cmpeq_epi8_mask(v, vq));
        other |= _mm512_cmpeq_epi8_mask(v, vc) | _mm512_cmpeq_epi8_mask(v, vn) | _mm512_cmpeq_epi8_mask(v, vr);
    }
    quotes = q; need = q != 0 || other != 0;
}
inline void csv_scan(const uint8_t* s, int64_t n, size_t& quotes, bool& need) {
    if (n >= 32 && wide()) csv_scan_wide(s, n, quotes, need); else csv_scan_scalar(s, n, quotes, need);
}

// ---- copy with every '"' doubled.  The wide form may store up to 63 bytes past the end of what it produced:
// the destination must have that much slack. ---------------------------------------------------------------
inline char* csv_double_quotes_scalar(char* o, const uint8_t* s, int64_t n) {
    for (int64_t i = 0; i < n; ++i) { if (s[i] == '"') *o++ = '"'; *o++ = (char)s[i]; }
    return o;
}
DYD_AVX512 inline char* csv_double_quotes_wide(char* o, const uint8_t* s, int64_t n) {
    const __m256i vq = _mm256_set1_epi8('"');
    const __m512i hi = _mm512_set1_epi16(0x2200);                  // '"' in the upper byte of every 16-bit lane
    int64_t i = 0;
    for (; i + 32 <= n; i += 32) {
        const __m256i x = _mm256_loadu_si256((const __m256i*)(s + i));
        const uint32_t q = _mm256_cmpeq_epi8_mask(x, vq);
        if (!q) { _mm256_storeu_si256((__m256i*)o, x); o += 32; continue; }
        // byte k -> slots 2k (the byte) and 2k+1 (a quote, kept only behind a quote), then squeeze the kept slots together
        const __m512i w = _mm512_or_si512(_mm512_cvtepu8_epi16(x), hi);
        const uint64_t keep = 0x5555555555555555ULL | _pdep_u64(q, 0xAAAAAAAAAAAAAAAAULL);
        _mm512_storeu_si512(o, _mm512_maskz_compress_epi8(keep, w));
        o += 32 + _mm_popcnt_u32(q);
    }
    if (i < n) {
        const uint32_t m = (~0u) >> (32 - (n - i));
        const __m256i x = _mm256_maskz_loadu_epi8(m, s + i);
        const uint32_t q = _mm256_cmpeq_epi8_mask(x, vq) & m;
        const __m512i w = _mm512_or_si512(_mm512_cvtepu8_epi16(x), hi);
        const uint64_t keep = _pdep_u64(m, 0x5555555555555555ULL) | _pdep_u64(q, 0xAAAAAAAAAAAAAAAAULL);
        const int cnt = (int)(n - i) + _mm_popcnt_u32(q);
        _mm512_mask_storeu_epi8(o, (~0ULL) >> (64 - cnt), _mm512_maskz_compress_epi8(keep, w));
        o += cnt;
    }
    return o;
}
// `slack` = the caller guarantees 64 writable bytes behind the result
inline char* csv_double_quotes(char* o, const uint8_t* s, int64_t n, bool slack) {
    if (slack && n >= 32 && wide()) return csv_double_quotes_wide(o, s, n);
    return csv_double_quotes_scalar(o, s, n);
}

// ---- any NUL byte in [s, s+n) ----------------------------------------------------------------------------
inline bool has_nul(const uint8_t* s, int64_t n) { return n > 0 && memchr(s, 0, (size_t)n) != nullptr; }

}  // namespace dyd_simd

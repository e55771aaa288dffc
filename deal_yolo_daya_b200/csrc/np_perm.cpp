// np.random.RandomState(seed).permutation(n), bit for bit, for the split step (processor.py:800: DataFrame.sample(frac=1,
// random_state=seed) draws exactly that permutation -- SURVEY.md 8a6) -- host C++.
//
// numpy's legacy generator: MT19937 seeded by init_genrand(seed); permutation(n) = arange(n) shuffled by
//   for i = n-1 .. 1:  j = random_interval(i);  swap(x[i], x[j])
// random_interval(max): mask = next power of two above max, minus one; draw 32-bit words (64-bit ones, two words high-first,
// when max > 0xffffffff) until (word & mask) <= max.  The sequence of (i, j) pairs does not depend on the array, so it is
// produced a block ahead and the swap targets are prefetched: numpy's own loop spends its time in the cache misses of x[j]
// (24 ns per element at 80 M elements, this one ~5).
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/dyd.h"

namespace {

struct MT {
    uint32_t key[624];
    int pos;
    explicit MT(uint32_t seed) {
        for (int i = 0; i < 624; ++i) { key[i] = seed; seed = 1812433253u * (seed ^ (seed >> 30)) + (uint32_t)i + 1u; }
        pos = 624;
    }
    void gen() {
        const uint32_t UP = 0x80000000u, LO = 0x7fffffffu, MAT = 0x9908b0dfu;
        int k = 0;
        for (; k < 624 - 397; ++k) { const uint32_t y = (key[k] & UP) | (key[k + 1] & LO); key[k] = key[k + 397] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MAT); }
        for (; k < 623; ++k) { const uint32_t y = (key[k] & UP) | (key[k + 1] & LO); key[k] = key[k + (397 - 624)] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MAT); }
        const uint32_t y = (key[623] & UP) | (key[0] & LO);
        key[623] = key[396] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MAT);
        pos = 0;
    }
    inline uint32_t next32() {
        if (pos == 624) gen();
        uint32_t y = key[pos++];
        y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
        return y;
    }
    inline uint64_t next64() { const uint64_t hi = next32(); return (hi << 32) | next32(); }
};

}  // namespace

extern "C" int dyd_numpy_permutation(uint32_t seed, int64_t n, int64_t* out) {
    if (n < 0 || (n > 0 && !out)) return DYD_E_ARG;
    for (int64_t i = 0; i < n; ++i) out[i] = i;
    if (n < 2) return 0;
    MT mt(seed);
    constexpr int B = 256;                                   // swap targets known this many steps ahead
    int64_t js[2][B + 1];
    uint32_t words[624];                                     // tempered outputs of one generator refill
    int wp = 624;
    auto refill = [&] {
        mt.gen();
        for (int k = 0; k < 624; ++k) {
            uint32_t y = mt.key[k];
            y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
            words[k] = y;
        }
        mt.pos = 624; wp = 0;
    };
    // targets of steps `from`, from-1, ... (at most B of them): random_interval(i) = first masked draw <= i.  The mask only
    // changes when i crosses a power of two, so inside such a stretch the accept / reject decision is applied without a branch.
    auto fill = [&](int which, int64_t from) {
        int c = 0;
        int64_t i = from;
        int64_t* dst = js[which];
        while (i >= 1 && c < B) {
            uint64_t mask = (uint64_t)i;
            mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16; mask |= mask >> 32;
            const int64_t lo = (int64_t)(mask >> 1);         // the mask holds for lo < i <= mask
            if ((uint64_t)i <= 0xffffffffULL) {
                const uint32_t m32 = (uint32_t)mask;
                while (i > lo && c < B) {
                    if (wp == 624) refill();
                    const int64_t v = (int64_t)(words[wp++] & m32);
                    dst[c] = v;
                    const int ok = v <= i;
                    c += ok; i -= ok;
                }
            } else {
                while (i > lo && c < B) {
                    if (wp == 624) refill();
                    const uint64_t hi = words[wp++];
                    if (wp == 624) refill();
                    const int64_t v = (int64_t)(((hi << 32) | words[wp++]) & mask);
                    dst[c] = v;
                    const int ok = v <= i;
                    c += ok; i -= ok;
                }
            }
        }
        for (int k = 0; k < c; ++k) __builtin_prefetch(out + dst[k], 1);
        return c;
    };
    int64_t i = n - 1;
    int cur = 0;
    int have = fill(cur, i);
    while (i >= 1) {
        const int nxt = fill(cur ^ 1, i - have);            // the block after this one is drawn (and prefetched) before this one is applied
        const int64_t* j = js[cur];
        for (int c = 0; c < have; ++c, --i) {
            const int64_t t = out[i]; out[i] = out[j[c]]; out[j[c]] = t;
        }
        cur ^= 1; have = nxt;
    }
    return 0;
}

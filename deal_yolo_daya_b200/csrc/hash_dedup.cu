// K0 string hash, K4 first/last-occurrence dedup, K5 anti-join -- device hash tables.
//
// Table: open addressing, linear probing, 16-byte slots {key, row}, capacity = 2^k >= 2n.
// One cudaMemset(0xFF) makes every key EMPTY and every row "no row yet" for both orders
// (UINT64_MAX under atomicMin, -1 under signed atomicMax).  A key equal to the EMPTY sentinel
// is folded onto EMPTY-1; like any 64-bit collision it is caught by the host's string check of
// dropped rows (d_rep / d_ref_row exist for that).
//
// K4 on large inputs does not touch that table at all: records are first scattered into 2^k
// partitions by the top bits of the mixed key (fixed-capacity regions, one atomic cursor each), then
// one CTA per partition deduplicates its few hundred records in a shared-memory table and writes
// keep / rep.  Random accesses hit shared memory instead of a DRAM-sized table; the global-table
// kernels remain as the path for small inputs and, gated by a device flag, for inputs that overflow
// a partition (one key repeated hundreds of times).
#include <stdlib.h>

#include <algorithm>
#include <type_traits>

#include "common.cuh"

namespace dyd {

constexpr unsigned long long EMPTY = 0xFFFFFFFFFFFFFFFFULL;
constexpr int HT_THREADS = 256;

struct Slot {
    unsigned long long key;
    unsigned long long row;
};
struct TableHeader {                 // 64 bytes in front of the slots
    unsigned long long null_first;   // min row among null cells (UINT64_MAX if none)
    long long null_last;             // max row among null cells (-1 if none)
    unsigned long long null_count;
    unsigned long long pad[5];
};

static inline uint64_t table_capacity(int64_t n) {
    uint64_t cap = 1024;
    while (cap < (uint64_t)(n > 0 ? n : 0) * 2) cap <<= 1;
    return cap;
}
static inline int log2u(uint64_t v) { int k = 0; while ((1ULL << k) < v) ++k; return k; }

__device__ __forceinline__ unsigned long long norm_key(unsigned long long k) { return k == EMPTY ? EMPTY - 1 : k; }
__device__ __forceinline__ uint64_t home_slot(unsigned long long k, int shift) { return (k * 0x9E3779B97F4A7C15ULL) >> shift; }

// ------------------------------------------------------------------------------- K0
// MurmurHash64A-style: 8-byte little-endian words, multiply-xorshift mixing.
constexpr unsigned long long HM = 0xC6A4A7935BD1E995ULL;
constexpr unsigned long long HSEED = 0x8445D61A4E774912ULL;

__global__ void __launch_bounds__(HT_THREADS)
hash_strings_kernel(const int64_t* __restrict__ off, const uint8_t* __restrict__ bytes, int64_t n,
                    uint64_t* __restrict__ out) {
    const int64_t r = blockIdx.x * (int64_t)HT_THREADS + threadIdx.x;
    if (r >= n) return;
    const int64_t a = off[r], len = off[r + 1] - a;
    unsigned long long h = HSEED ^ ((unsigned long long)len * HM);
    const int64_t nblk = len >> 3;
    const uint8_t* p = bytes + a;
    const unsigned sh = ((uintptr_t)p & 7) * 8;
    const unsigned long long* w = reinterpret_cast<const unsigned long long*>((uintptr_t)p & ~(uintptr_t)7);
    if (nblk > 0) {
        unsigned long long lo = __ldg(w);
        for (int64_t i = 0; i < nblk; ++i) {
            unsigned long long k;
            if (sh == 0) { k = lo; lo = (i + 1 < nblk) ? __ldg(w + i + 1) : 0ULL; }
            else { unsigned long long hi = __ldg(w + i + 1); k = (lo >> sh) | (hi << (64 - sh)); lo = hi; }
            k *= HM; k ^= k >> 47; k *= HM;
            h ^= k; h *= HM;
        }
    }
    const int rem = (int)(len & 7);
    if (rem) {
        unsigned long long t = 0;
        const uint8_t* q = p + (nblk << 3);
        for (int i = 0; i < rem; ++i) t |= (unsigned long long)__ldg(q + i) << (8 * i);
        h ^= t; h *= HM;
    }
    h ^= h >> 47; h *= HM; h ^= h >> 47;
    out[r] = h;
}

// ------------------------------------------------------------------------------- K4
// KIND 0: rows are 0..n-1 (null flags honoured); 1: explicit global row ids in a separate array;
// 2: interleaved (key, id) records as they arrive from the exchange (keys = records, row_id unused)
template <int KIND>
__global__ void __launch_bounds__(HT_THREADS)
dedup_insert_kernel(const unsigned long long* __restrict__ keys, const uint8_t* __restrict__ null,
                    const int64_t* __restrict__ row_id, int64_t n, int keep_mode,
                    TableHeader* hdr, Slot* tab, unsigned* cnt, int shift, uint64_t mask, int pass_shift, unsigned pass,
                    const int* gate, bool count_nulls) {
    if (gate != nullptr && *gate == 0) return;         // fallback path of the partitioned dedup: not needed
    for (int64_t r0 = blockIdx.x * (int64_t)HT_THREADS; r0 < n; r0 += (int64_t)gridDim.x * HT_THREADS) {
    const int64_t r = r0 + threadIdx.x;
    const bool live = r < n;
    constexpr bool IDS = KIND != 0;
    const bool isnull = live && !IDS && null != nullptr && null[r] != 0;
    const unsigned nm = (pass == 0 && count_nulls) ? __ballot_sync(FULL, isnull) : 0u;
    if (nm) {                                          // rows ascend with the lane: aggregate per warp
        const int lane = threadIdx.x & 31;
        if (lane == __ffs(nm) - 1) { atomicMin(&hdr->null_first, (unsigned long long)r); atomicAdd(&hdr->null_count, (unsigned long long)__popc(nm)); }
        if (lane == 31 - __clz(nm)) atomicMax(&hdr->null_last, (long long)r);
    }
    if (!live || isnull) continue;
    const long long id = KIND == 2 ? (long long)keys[2 * r + 1] : (KIND == 1 ? row_id[r] : r);
    if (IDS && id < 0) continue;                        // padding of a fixed-capacity exchange bucket
    const unsigned long long key = norm_key(KIND == 2 ? keys[2 * r] : keys[r]);
    const unsigned long long rid = (unsigned long long)id;
    uint64_t s = home_slot(key, shift);
    if ((unsigned)(s >> pass_shift) != pass) continue; // this pass works on another region of the table
    for (;;) {
        unsigned long long prev = atomicCAS(&tab[s].key, EMPTY, key);
        if (prev == EMPTY || prev == key) {
            if (keep_mode == 1) atomicMax(reinterpret_cast<long long*>(&tab[s].row), (long long)rid);
            else atomicMin(&tab[s].row, rid);
            if (keep_mode == 2) atomicAdd(&cnt[s], 1u);
            break;
        }
        s = (s + 1) & mask;
    }
    }
}

template <int KIND>
__global__ void __launch_bounds__(HT_THREADS)
dedup_lookup_kernel(const unsigned long long* __restrict__ keys, const uint8_t* __restrict__ null,
                    const int64_t* __restrict__ row_id, int64_t n, int keep_mode,
                    const TableHeader* __restrict__ hdr, const Slot* __restrict__ tab,
                    const unsigned* __restrict__ cnt, int shift, uint64_t mask, int pass_shift, unsigned pass,
                    uint8_t* __restrict__ keep, int64_t* __restrict__ rep, const int* gate) {
    if (gate != nullptr && *gate == 0) return;
    for (int64_t r = blockIdx.x * (int64_t)HT_THREADS + threadIdx.x; r < n; r += (int64_t)gridDim.x * HT_THREADS) {
    constexpr bool IDS = KIND != 0;
    const long long rid = KIND == 2 ? (long long)keys[2 * r + 1] : (KIND == 1 ? row_id[r] : r);
    if (IDS && rid < 0) { if (pass == 0) { rep[r] = -1; keep[r] = 0; } continue; }
    if (!IDS && null != nullptr && null[r] != 0) {
        if (pass != 0) continue;
        const long long rp = keep_mode == 1 ? hdr->null_last : (long long)hdr->null_first;
        rep[r] = rp;
        keep[r] = keep_mode == 2 ? (hdr->null_count == 1) : (rp == rid);
        continue;
    }
    const unsigned long long key = norm_key(KIND == 2 ? keys[2 * r] : keys[r]);
    uint64_t s = home_slot(key, shift);
    if ((unsigned)(s >> pass_shift) != pass) continue;
    while (tab[s].key != key) s = (s + 1) & mask;      // the key was inserted by the previous kernel
    const long long rp = (long long)tab[s].row;
    rep[r] = rp;
    keep[r] = keep_mode == 2 ? (cnt[s] == 1u) : (rp == rid);
    }
}

// ------------------------------------------------------------------------------- K4, partitioned
#ifndef DYD_PT_THREADS
#define DYD_PT_THREADS 128
#endif
#ifndef DYD_PT_FILL
#define DYD_PT_FILL 128                          // average records per partition in [FILL, 2 * FILL)
#endif
constexpr int PT_SLOTS = 1024;                   // shared-memory table of one partition
constexpr int PT_THREADS = DYD_PT_THREADS;
constexpr int PT_MAX_PER_THREAD = 512 / PT_THREADS;   // partition capacity < PT_THREADS * PT_MAX_PER_THREAD = 512
constexpr unsigned long long GOLD = 0x9E3779B97F4A7C15ULL;

struct PartLayout {                              // carved out of the caller's workspace
    int log2_np;
    unsigned pcap;                               // records per partition region (2x the average fill)
    size_t cursors, flag, kv, rr, fallback, total;   // byte offsets from the workspace start
};
static inline PartLayout part_layout(int64_t n, bool with_r) {
    PartLayout L{};
    int k = 0;
    while (((long long)DYD_PT_FILL << (k + 1)) <= n) ++k;   // 2^k partitions, average fill in [FILL, 2 * FILL)
    L.log2_np = k;
    const uint64_t np = 1ULL << k;
    L.pcap = (unsigned)((2 * (uint64_t)n) / np);
    if (L.pcap >= (unsigned)(PT_THREADS * PT_MAX_PER_THREAD)) L.pcap = PT_THREADS * PT_MAX_PER_THREAD - 1;
    size_t o = sizeof(TableHeader);
    L.cursors = o; o += sizeof(unsigned) * np;
    L.flag = o; o += 16;
    o = (o + 15) & ~(size_t)15;
    L.kv = o; o += sizeof(ulonglong2) * np * L.pcap;
    L.rr = o; o += with_r ? sizeof(unsigned) * np * L.pcap : 0;
    o = (o + 255) & ~(size_t)255;
    L.fallback = o;                              // a complete global-table workspace for the gated fallback
    const uint64_t cap = table_capacity(n);
    L.total = o + sizeof(TableHeader) + cap * sizeof(Slot) + cap * sizeof(unsigned);
    return L;
}
// Partitioning pays once the table no longer sits in L2 (measured on B200: 1 M rows 0.074 vs 0.060 ms,
// 4 M rows 0.18 vs 0.24 ms, 10 M rows 0.45 vs 0.79 ms).  DYD_DEDUP_PARTITION=0 switches it off,
// DYD_DEDUP_PARTITION_MIN moves the threshold (the tests lower it to cover the path with small inputs).
static inline bool use_partitions(int64_t n) {
    const char* e = getenv("DYD_DEDUP_PARTITION");
    if (e && atoi(e) == 0) return false;
    const char* m = getenv("DYD_DEDUP_PARTITION_MIN");
    const long long lo = m ? atoll(m) : (1LL << 21);
    return n >= (lo < 65536 ? 65536 : lo) && n < (1LL << 31);
}

template <int KIND>
__global__ void __launch_bounds__(HT_THREADS)
dedup_partition_kernel(const unsigned long long* __restrict__ keys, const uint8_t* __restrict__ null,
                       const int64_t* __restrict__ row_id, int64_t n, TableHeader* hdr, unsigned* cursors, int* overflow,
                       ulonglong2* __restrict__ kv, unsigned* __restrict__ rr, int pshift, unsigned pcap,
                       uint8_t* __restrict__ keep, int64_t* __restrict__ rep,
                       uint8_t* __restrict__ keep2 = nullptr, int64_t* __restrict__ ref2 = nullptr) {
    const int64_t r = blockIdx.x * (int64_t)HT_THREADS + threadIdx.x;
    const bool live = r < n;
    constexpr bool IDS = KIND != 0;
    // joint form (dyd_url_filter): the anti-join answer of a row no reference key meets, written here with coalesced stores;
    // the resolve kernel rewrites only the rows it finds in the reference set.  Bucket padding answers 0 / -1.
    if (live && keep2 != nullptr) {
        keep2[r] = (KIND == 2 && (long long)keys[2 * r + 1] < 0) || (KIND == 1 && row_id[r] < 0) ? 0 : 1;
        ref2[r] = -1;
    }
    const bool isnull = live && !IDS && null != nullptr && null[r] != 0;
    const unsigned nm = __ballot_sync(FULL, isnull);
    if (nm) {                                          // rows ascend with the lane: aggregate per warp
        const int lane = threadIdx.x & 31;
        if (lane == __ffs(nm) - 1) { atomicMin(&hdr->null_first, (unsigned long long)r); atomicAdd(&hdr->null_count, (unsigned long long)__popc(nm)); }
        if (lane == 31 - __clz(nm)) atomicMax(&hdr->null_last, (long long)r);
    }
    if (!live || isnull) return;
    const long long id = KIND == 2 ? (long long)keys[2 * r + 1] : (KIND == 1 ? row_id[r] : r);
    if (IDS && id < 0) { rep[r] = -1; keep[r] = 0; return; }   // padding of a fixed-capacity exchange bucket: answered here
    const unsigned long long key = norm_key(KIND == 2 ? keys[2 * r] : keys[r]);
    const unsigned p = pshift >= 64 ? 0u : (unsigned)((key * GOLD) >> pshift);
    const unsigned slot = atomicAdd(&cursors[p], 1u);
    if (slot >= pcap) { *overflow = 1; return; }
    const size_t at = (size_t)p * pcap + slot;
    kv[at] = make_ulonglong2(key, (unsigned long long)id);
    if (IDS) rr[at] = (unsigned)r;
    // the answer of a row that turns out to be its group's survivor, written here with coalesced stores: the resolve
    // kernel then only touches the rows it drops (a few percent) instead of scattering 9 bytes to every row
    rep[r] = id; keep[r] = 1;
}

// One CTA per partition: shared-memory table, then keep / rep of every record of the partition.
// ROW32: row ids fit 32 bits (KIND 0, n < 2^31), so first / last are native 32-bit shared atomics.
template <int KIND, int MODE, bool ROW32>
__global__ void __launch_bounds__(PT_THREADS)
dedup_resolve_kernel(const unsigned* __restrict__ cursors, const int* __restrict__ overflow, const ulonglong2* __restrict__ kv,
                     const unsigned* __restrict__ rr, int pshift, unsigned pcap,
                     uint8_t* __restrict__ keep, int64_t* __restrict__ rep) {
    using RowT = typename std::conditional<ROW32, unsigned, unsigned long long>::type;
    using SRowT = typename std::conditional<ROW32, int, long long>::type;
    __shared__ unsigned long long skey[PT_SLOTS];
    __shared__ RowT srow[PT_SLOTS];
    __shared__ unsigned scnt[MODE == 2 ? PT_SLOTS : 1];
    if (*overflow) return;                             // the gated global-table kernels take over
    const unsigned p = blockIdx.x;
    const unsigned cnt = min(cursors[p], pcap);
    if (cnt == 0) return;
    for (int i = threadIdx.x; i < PT_SLOTS; i += PT_THREADS) {
        skey[i] = EMPTY; srow[i] = (RowT)~(RowT)0;     // "no row yet" for min (max unsigned) and for signed max (-1)
        if (MODE == 2) scnt[i] = 0;
    }
    __syncthreads();
    const ulonglong2* mine = kv + (size_t)p * pcap;
    unsigned long long id[PT_MAX_PER_THREAD];
    int at[PT_MAX_PER_THREAD];
#pragma unroll
    for (int u = 0; u < PT_MAX_PER_THREAD; ++u) {
        const unsigned i = threadIdx.x + u * PT_THREADS;
        at[u] = -1;
        if (i < cnt) {
            const ulonglong2 rec = mine[i];
            id[u] = rec.y;
            // bits just below the partition bits pick the home slot
            unsigned s = (unsigned)(((rec.x * GOLD) << (64 - pshift)) >> 54) & (PT_SLOTS - 1);
            for (;;) {
                const unsigned long long prev = atomicCAS(&skey[s], EMPTY, rec.x);
                if (prev == EMPTY || prev == rec.x) break;
                s = (s + 1) & (PT_SLOTS - 1);
            }
            if (MODE == 1) atomicMax(reinterpret_cast<SRowT*>(&srow[s]), (SRowT)rec.y);
            else atomicMin(&srow[s], (RowT)rec.y);
            if (MODE == 2) atomicAdd(&scnt[s], 1u);
            at[u] = (int)s;
        }
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < PT_MAX_PER_THREAD; ++u) {
        if (at[u] < 0) continue;
        const unsigned i = threadIdx.x + u * PT_THREADS;
        const long long rp = (long long)(SRowT)srow[at[u]];
        const bool kept = MODE == 2 ? (scnt[at[u]] == 1u) : (rp == (long long)id[u]);
        if (kept && rp == (long long)id[u]) continue;                  // already answered by the partition kernel
        const size_t out = KIND == 0 ? (size_t)id[u] : (size_t)rr[(size_t)p * pcap + i];
        rep[out] = rp;
        keep[out] = kept ? 1 : 0;
    }
}

// rows the partitions never saw: null cells of KIND 0 (bucket padding of KIND 1 / 2 is answered by the partition kernel)
template <int KIND>
__global__ void __launch_bounds__(HT_THREADS)
dedup_leftover_kernel(const unsigned long long* __restrict__ keys, const uint8_t* __restrict__ null,
                      const int64_t* __restrict__ row_id, int64_t n, int keep_mode, const TableHeader* __restrict__ hdr,
                      const int* __restrict__ overflow, uint8_t* __restrict__ keep, int64_t* __restrict__ rep) {
    if (*overflow) return;
    const int64_t r = blockIdx.x * (int64_t)HT_THREADS + threadIdx.x;
    if (r >= n) return;
    if (KIND == 0) {
        if (null == nullptr || null[r] == 0) return;
        const long long rp = keep_mode == 1 ? hdr->null_last : (long long)hdr->null_first;
        rep[r] = rp;
        keep[r] = keep_mode == 2 ? (hdr->null_count == 1) : (rp == r);
    } else {
        const long long id = KIND == 2 ? (long long)keys[2 * r + 1] : row_id[r];
        if (id < 0) { rep[r] = -1; keep[r] = 0; }
    }
}

// ------------------------------------------------------------------------------- K5
// KIND 0: plain key arrays (row = index, null flags honoured); KIND 2: interleaved (key, id) records as they arrive
// from the exchange (id < 0 = padding of a fixed-capacity bucket; `null` unused).  With `reset` the build kernel,
// the last reader of the reference records, turns every record back into padding for the next exchange step.
template <int KIND>
__global__ void __launch_bounds__(HT_THREADS)
antijoin_build_kernel(unsigned long long* __restrict__ keys, const uint8_t* __restrict__ null, int64_t n,
                      Slot* tab, int shift, uint64_t mask, bool reset, const int* gate) {
    if (gate != nullptr && *gate == 0) return;         // fallback of the partitioned path: not needed
    const int64_t r = blockIdx.x * (int64_t)HT_THREADS + threadIdx.x;
    if (r >= n) return;
    unsigned long long key, row;
    if (KIND == 0) {
        if (null != nullptr && null[r] != 0) return;               // ref.dropna()
        key = norm_key(keys[r]); row = (unsigned long long)r;
    } else {
        const ulonglong2 rec = reinterpret_cast<const ulonglong2*>(keys)[r];
        if ((long long)rec.y < 0) return;
        key = norm_key(rec.x); row = rec.y;
        if (reset) keys[2 * r + 1] = ~0ULL;
    }
    uint64_t s = home_slot(key, shift);
    for (;;) {
        unsigned long long prev = atomicCAS(&tab[s].key, EMPTY, key);
        if (prev == EMPTY || prev == key) { atomicMin(&tab[s].row, row); return; }
        s = (s + 1) & mask;
    }
}

template <int KIND>
__global__ void __launch_bounds__(HT_THREADS)
antijoin_probe_kernel(const unsigned long long* __restrict__ keys, const uint8_t* __restrict__ null, int64_t n,
                      const Slot* __restrict__ tab, int shift, uint64_t mask,
                      uint8_t* __restrict__ keep, int64_t* __restrict__ ref_row, const int* gate) {
    if (gate != nullptr && *gate == 0) return;
    const int64_t r = blockIdx.x * (int64_t)HT_THREADS + threadIdx.x;
    if (r >= n) return;
    uint8_t k = 1; long long rr = -1;
    bool live;
    unsigned long long key;
    if (KIND == 0) { live = null == nullptr || null[r] == 0; key = live ? keys[r] : 0ULL; }   // a NaN main cell never matches
    else { const ulonglong2 rec = reinterpret_cast<const ulonglong2*>(keys)[r]; live = (long long)rec.y >= 0; key = rec.x; if (!live) k = 0; }
    if (live) {
        key = norm_key(key);
        uint64_t s = home_slot(key, shift);
        for (;;) {
            const unsigned long long cur = tab[s].key;
            if (cur == key) { k = 0; rr = (long long)tab[s].row; break; }
            if (cur == EMPTY) break;
            s = (s + 1) & mask;
        }
    }
    keep[r] = k; ref_row[r] = rr;
}

// ------------------------------------------------------------------------------- multi-GPU exchange helpers
// Owner rank of a key; must match sharding.owner_of() on the Python side.
__device__ __forceinline__ int owner_of(unsigned long long key, int world) {
    const unsigned long long mixed = key * 0x9E3779B97F4A7C15ULL;
    return (int)(((mixed >> 33) & 0x7FFFFFFFULL) % (unsigned)world);
}

// (key, global row id) records scattered into `world` fixed-capacity buckets (pre-filled with padding
// by a 0xFF memset: key = EMPTY, id = -1).  Order inside a bucket is arbitrary: the ids carry it.
__global__ void __launch_bounds__(HT_THREADS)
shard_bucket_kernel(const unsigned long long* __restrict__ keys, const uint8_t* __restrict__ null, int64_t row_base,
                    int64_t n, int world, int64_t cap, long long* __restrict__ records,
                    unsigned long long* cursors, int* overflow) {
    const int64_t r = blockIdx.x * (int64_t)HT_THREADS + threadIdx.x;
    const bool live = r < n && (null == nullptr || null[r] == 0);
    const unsigned long long key = live ? keys[r] : 0ULL;
    const int own = live ? owner_of(key, world) : -1;
    const unsigned peers = __match_any_sync(FULL, own);          // one atomic per (warp, owner)
    if (!live) return;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(peers) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(&cursors[own], (unsigned long long)__popc(peers));
    base = __shfl_sync(peers, base, leader);
    const unsigned long long slot = base + __popc(peers & ((1u << lane) - 1u));
    if (slot >= (unsigned long long)cap) { *overflow = 1; return; }
    long long* rec = records + 2 * ((int64_t)own * cap + (int64_t)slot);
    rec[0] = (long long)key; rec[1] = row_base + r;
}

// The same scatter straight into the owners' receive buffers over NVLink peer memory: rank `me`
// owns region `me` of every peer's buffer, so the bucket step IS the all-to-all (no separate collective).
constexpr int P2P_MAX_WORLD = 64;                 // block-aggregated cursors up to this many ranks
__global__ void __launch_bounds__(HT_THREADS)
shard_bucket_p2p_kernel(const unsigned long long* __restrict__ keys, const uint8_t* __restrict__ null, int64_t row_base,
                        int64_t n, int world, int me, int64_t cap, long long* const* __restrict__ peer_records,
                        unsigned* __restrict__ sent_row, unsigned long long* cursors, int* overflow) {
    // One global atomic per (block, owner): the block counts its records per owner in shared memory,
    // claims a contiguous run of slots in each owner's region and fills it -- runs of ~256 / world
    // records (16 bytes each) keep the NVLink stores wide and the cursor traffic negligible.
    __shared__ unsigned scnt[P2P_MAX_WORLD];
    __shared__ unsigned long long sbase[P2P_MAX_WORLD];
    const int64_t r = blockIdx.x * (int64_t)HT_THREADS + threadIdx.x;
    const bool live = r < n && (null == nullptr || null[r] == 0);
    const unsigned long long key = live ? keys[r] : 0ULL;
    const int own = live ? owner_of(key, world) : -1;
    unsigned long long slot;
    if (world <= P2P_MAX_WORLD) {
        if (threadIdx.x < world) scnt[threadIdx.x] = 0;
        __syncthreads();
        const unsigned rank_in_block = live ? atomicAdd(&scnt[own], 1u) : 0u;
        __syncthreads();
        if (threadIdx.x < world && scnt[threadIdx.x]) sbase[threadIdx.x] = atomicAdd(&cursors[threadIdx.x], (unsigned long long)scnt[threadIdx.x]);
        __syncthreads();
        if (!live) return;
        slot = sbase[own] + rank_in_block;
    } else {                                           // one atomic per (warp, owner)
        const unsigned peers = __match_any_sync(FULL, own);
        if (!live) return;
        const int lane = threadIdx.x & 31;
        const int leader = __ffs(peers) - 1;
        unsigned long long base = 0;
        if (lane == leader) base = atomicAdd(&cursors[own], (unsigned long long)__popc(peers));
        base = __shfl_sync(peers, base, leader);
        slot = base + __popc(peers & ((1u << lane) - 1u));
    }
    if (slot >= (unsigned long long)cap) { *overflow = 1; return; }
    longlong2* rec = reinterpret_cast<longlong2*>(peer_records[own]) + ((int64_t)me * cap + (int64_t)slot);
    *rec = make_longlong2((long long)key, row_base + r);           // one 16-byte store per record
    if (sent_row != nullptr) sent_row[(int64_t)own * cap + (int64_t)slot] = (unsigned)r;   // local: which row the answer in that slot belongs to
}

// owner side, peer-memory form: the answer for a record that came from rank s goes straight into
// region `me` of rank s's reply buffer
// mode 0 (dedup): answer = rep | keep << 62; mode 1 (anti-join): answer = kept ? 1 << 62 : first matching reference row.
// This kernel is the last reader of the received records: with `reset` it turns each one back into padding (id = -1),
// so the next step needs no separate fill of the receive buffer and no "buffers are clean" barrier.
__device__ __forceinline__ long long reply_word(int mode, long long id, uint8_t keep, long long rep) {
    if (id < 0) return -1;
    if (mode == 1) return keep ? (1LL << 62) : rep;
    return rep | ((long long)(keep ? 1 : 0) << 62);
}
__global__ void __launch_bounds__(HT_THREADS)
shard_pack_reply_p2p_kernel(long long* __restrict__ records, const uint8_t* __restrict__ keep,
                            const int64_t* __restrict__ rep, int64_t m, int64_t cap, int me,
                            long long* const* __restrict__ peer_reply, int mode, bool reset) {
    const int64_t r = blockIdx.x * (int64_t)HT_THREADS + threadIdx.x;
    if (r >= m) return;
    const int src = (int)(r / cap);
    const int64_t slot = r - (int64_t)src * cap;
    const long long id = records[2 * r + 1];
    if (reset && id >= 0) records[2 * r + 1] = -1;
    // 8 bytes per answer: the origin remembers which of its rows sits in (owner, slot)
    peer_reply[src][(int64_t)me * cap + slot] = reply_word(mode, id, keep[r], rep[r]);
}

// origin side, peer-memory form: slot s of owner o's region holds the answer for row sent_row[o][s]
__global__ void __launch_bounds__(HT_THREADS)
shard_unpack_p2p_kernel(const long long* __restrict__ reply, const unsigned* __restrict__ sent_row,
                        const unsigned long long* __restrict__ cursors, int world, int64_t cap, int64_t n,
                        uint8_t* __restrict__ keep, int64_t* __restrict__ rep, int mode) {
    const int64_t t = blockIdx.x * (int64_t)HT_THREADS + threadIdx.x;
    if (t >= (int64_t)world * cap) return;
    const int own = (int)(t / cap);
    const int64_t slot = t - (int64_t)own * cap;
    if ((unsigned long long)slot >= cursors[own]) return;           // never filled
    const long long v = reply[t];
    const int64_t row = sent_row[t];
    if (v < 0 || row >= n) return;
    const uint8_t k = (uint8_t)((v >> 62) & 1);
    keep[row] = k;
    rep[row] = (mode == 1 && k) ? -1 : (v & ((1LL << 62) - 1));
}

// Staged form of the same scatter (world <= P2P_MAX_WORLD): a block takes ST_PER_BLOCK records, counting-sorts them by
// owner in shared memory, claims one run of slots per owner and writes each run with consecutive threads -> consecutive
// 16-byte stores, i.e. full-width NVLink packets instead of the 16 .. 64-byte pieces of the per-thread form.
constexpr int ST_THREADS = 512;
constexpr int ST_PER_THREAD = 4;
constexpr int ST_PER_BLOCK = ST_THREADS * ST_PER_THREAD;
__global__ void __launch_bounds__(ST_THREADS)
shard_bucket_staged_kernel(const unsigned long long* __restrict__ keys, const uint8_t* __restrict__ null, int64_t row_base,
                           int64_t n, int world, int me, int64_t cap, long long* const* __restrict__ peer_records,
                           unsigned* __restrict__ sent_row, unsigned long long* cursors, int* overflow,
                           uint8_t* __restrict__ def_keep_d, int64_t* __restrict__ def_rep_d, uint8_t* __restrict__ def_keep_a,
                           int64_t* __restrict__ def_ref_row) {
    __shared__ unsigned long long skey[ST_PER_BLOCK];
    __shared__ unsigned srow[ST_PER_BLOCK];
    __shared__ unsigned scnt[P2P_MAX_WORLD], soff[P2P_MAX_WORLD + 1];
    __shared__ unsigned long long sbase[P2P_MAX_WORLD];
    const int64_t r0 = blockIdx.x * (int64_t)ST_PER_BLOCK;
    if (threadIdx.x < world) scnt[threadIdx.x] = 0;
    __syncthreads();
    unsigned long long key[ST_PER_THREAD];
    int own[ST_PER_THREAD];
    unsigned rank[ST_PER_THREAD];
#pragma unroll
    for (int u = 0; u < ST_PER_THREAD; ++u) {
        const int64_t r = r0 + u * ST_THREADS + threadIdx.x;
        const bool live = r < n && (null == nullptr || null[r] == 0);
        key[u] = live ? keys[r] : 0ULL;
        own[u] = live ? owner_of(key[u], world) : -1;
        rank[u] = live ? atomicAdd(&scnt[own[u]], 1u) : 0u;
        // sparse replies (sharding.UrlFilterExchange): the answer of a row that survives both questions is written here,
        // coalesced; the owners then send back -- and the unpack kernel scatters -- only the few per cent that differ
        if (def_keep_d != nullptr && r < n) { def_keep_d[r] = 1; def_rep_d[r] = row_base + r; def_keep_a[r] = 1; def_ref_row[r] = -1; }
    }
    __syncthreads();
    if (threadIdx.x == 0) { unsigned acc = 0; for (int o = 0; o < world; ++o) { soff[o] = acc; acc += scnt[o]; } soff[world] = acc; }
    if (threadIdx.x < world && scnt[threadIdx.x]) sbase[threadIdx.x] = atomicAdd(&cursors[threadIdx.x], (unsigned long long)scnt[threadIdx.x]);
    __syncthreads();
#pragma unroll
    for (int u = 0; u < ST_PER_THREAD; ++u) {
        if (own[u] < 0) continue;
        const unsigned at = soff[own[u]] + rank[u];
        skey[at] = key[u];
        srow[at] = (unsigned)(u * ST_THREADS + threadIdx.x);                 // row inside the block
    }
    __syncthreads();
    const unsigned total = soff[world];
    for (unsigned p = threadIdx.x; p < total; p += ST_THREADS) {
        int o = 0;
        while (soff[o + 1] <= p) ++o;                                     // world <= 64: a short walk, the same for neighbouring threads
        const unsigned long long slot = sbase[o] + (p - soff[o]);
        if (slot >= (unsigned long long)cap) { *overflow = 1; continue; }
        const int64_t r = r0 + srow[p];
        longlong2* rec = reinterpret_cast<longlong2*>(peer_records[o]) + ((int64_t)me * cap + (int64_t)slot);
        *rec = make_longlong2((long long)skey[p], row_base + r);
        if (sent_row != nullptr) sent_row[(int64_t)o * cap + (int64_t)slot] = (unsigned)r;
    }
}

// Both answers of the joint exchange (sharding.UrlFilterExchange) in one 16-byte store per record: word 0 = dedup
// (rep | keep << 62), word 1 = anti-join (kept ? 1 << 62 : first matching reference row); padding records answer (-1, -1).
__global__ void __launch_bounds__(HT_THREADS)
shard_pack_reply2_p2p_kernel(long long* __restrict__ records, const uint8_t* __restrict__ keep_d, const int64_t* __restrict__ rep_d,
                             const uint8_t* __restrict__ keep_a, const int64_t* __restrict__ rep_a, int64_t m, int64_t cap, int me,
                             long long* const* __restrict__ peer_reply, bool reset, bool sparse) {
    const int64_t r = blockIdx.x * (int64_t)HT_THREADS + threadIdx.x;
    if (r >= m) return;
    const int src = (int)(r / cap);
    const int64_t slot = r - (int64_t)src * cap;
    const long long id = records[2 * r + 1];
    if (reset && id >= 0) records[2 * r + 1] = -1;
    // sparse: the origin already holds "kept by both" for every row and its reply buffer says "no answer" everywhere
    if (sparse && (id < 0 || (keep_d[r] != 0 && rep_d[r] == id && keep_a[r] != 0))) return;
    reinterpret_cast<longlong2*>(peer_reply[src])[(int64_t)me * cap + slot] =
        make_longlong2(reply_word(0, id, keep_d[r], rep_d[r]), reply_word(1, id, keep_a[r], rep_a[r]));
}
__global__ void __launch_bounds__(HT_THREADS)
shard_unpack2_p2p_kernel(longlong2* __restrict__ reply, const unsigned* __restrict__ sent_row,
                         const unsigned long long* __restrict__ cursors, int world, int64_t cap, int64_t n,
                         uint8_t* __restrict__ keep_d, int64_t* __restrict__ rep_d, uint8_t* __restrict__ keep_a, int64_t* __restrict__ rep_a,
                         bool sparse) {
    const int64_t t = blockIdx.x * (int64_t)HT_THREADS + threadIdx.x;
    if (t >= (int64_t)world * cap) return;
    const int own = (int)(t / cap);
    const int64_t slot = t - (int64_t)own * cap;
    if ((unsigned long long)slot >= cursors[own]) return;           // never filled
    const longlong2 v = reply[t];
    if (v.x < 0) return;                                            // padding, or (sparse) nothing to change
    if (sparse) reply[t] = make_longlong2(-1, -1);                  // last reader: the slot says "no answer" again for the next step
    const int64_t row = sent_row[t];
    if (row >= n) return;
    keep_d[row] = (uint8_t)((v.x >> 62) & 1);
    rep_d[row] = v.x & ((1LL << 62) - 1);
    const uint8_t k = (uint8_t)((v.y >> 62) & 1);
    keep_a[row] = k;
    rep_a[row] = k ? -1 : (v.y & ((1LL << 62) - 1));
}

// owner side: (id, rep | keep << 62) per received record
__global__ void __launch_bounds__(HT_THREADS)
shard_pack_reply_kernel(const long long* __restrict__ records, const uint8_t* __restrict__ keep,
                        const int64_t* __restrict__ rep, int64_t m, long long* __restrict__ reply, int mode) {
    const int64_t r = blockIdx.x * (int64_t)HT_THREADS + threadIdx.x;
    if (r >= m) return;
    const long long id = records[2 * r + 1];
    reply[2 * r] = id;
    reply[2 * r + 1] = reply_word(mode, id, keep[r], rep[r]);
}

// origin side: place the answers at the rows they belong to
__global__ void __launch_bounds__(HT_THREADS)
shard_unpack_kernel(const long long* __restrict__ reply, int64_t m, int64_t row_base, int64_t n,
                    uint8_t* __restrict__ keep, int64_t* __restrict__ rep, int mode) {
    const int64_t r = blockIdx.x * (int64_t)HT_THREADS + threadIdx.x;
    if (r >= m) return;
    const long long id = reply[2 * r];
    if (id < 0) return;
    const long long local = id - row_base;
    if (local < 0 || local >= n) return;
    const long long v = reply[2 * r + 1];
    const uint8_t k = (uint8_t)((v >> 62) & 1);
    keep[local] = k;
    rep[local] = (mode == 1 && k) ? -1 : (v & ((1LL << 62) - 1));
}

static inline unsigned grid_for(int64_t n) { return (unsigned)((n + HT_THREADS - 1) / HT_THREADS); }
// EXPERIMENT: dynamic shared memory requested by every URL-chain launch.  A non-zero value keeps these kernels off the SMs
// whose shared memory the persistent fused kernel owns, i.e. confines them to the SMs that kernel leaves free.
static inline size_t url_pad() { const char* e = getenv("DYD_URL_SMEM_PAD"); return e ? (size_t)atoi(e) : 0; }

// memset that only happens when *gate != 0 (16-byte units)
__global__ void gated_fill_kernel(uint4* p, size_t n16, unsigned v, const int* gate) {
    if (*gate == 0) return;
    const uint4 x = make_uint4(v, v, v, v);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) p[i] = x;
}

// the global-table path; gate == nullptr: unconditional, else only if *gate != 0
template <int KIND>
static int dedup_table(const uint64_t* d_keys, const uint8_t* d_null, const int64_t* d_row_id, int64_t n, int keep_mode,
                       uint8_t* d_keep, int64_t* d_rep, void* ws, const int* gate, TableHeader* null_hdr, cudaStream_t s) {
    const uint64_t cap = table_capacity(n);
    const int shift = 64 - log2u(cap);
    TableHeader* hdr = reinterpret_cast<TableHeader*>(ws);
    Slot* tab = reinterpret_cast<Slot*>(hdr + 1);
    unsigned* cnt = reinterpret_cast<unsigned*>(tab + cap);
    if (gate == nullptr) {
        DYD_CUDA(cudaMemsetAsync(ws, 0xFF, sizeof(TableHeader) + cap * sizeof(Slot), s));
        DYD_CUDA(cudaMemsetAsync(&hdr->null_count, 0, sizeof(unsigned long long), s));
        if (keep_mode == 2) DYD_CUDA(cudaMemsetAsync(cnt, 0, cap * sizeof(unsigned), s));
    } else {
        gated_fill_kernel<<<NUM_SMS * 8, 256, url_pad(), s>>>(reinterpret_cast<uint4*>(tab), cap * sizeof(Slot) / 16, 0xFFFFFFFFu, gate);
        if (int rc = launch_check("gated_fill_kernel")) return rc;
        if (keep_mode == 2) {
            gated_fill_kernel<<<NUM_SMS * 8, 256, url_pad(), s>>>(reinterpret_cast<uint4*>(cnt), cap * sizeof(unsigned) / 16, 0u, gate);
            if (int rc = launch_check("gated_fill_kernel")) return rc;
        }
        hdr = null_hdr;                                // null statistics were gathered by the partition kernel
    }
    // Large tables are worked region by region: each pass touches 1/2^k of the table (a slice that
    // stays L2-resident) and skips the keys that hash elsewhere; the key array itself streams.
    // (measured on B200, 10 M keys / 537 MB table: 1 pass 0.94 ms, 2 passes 0.79 ms, 4 passes 0.93 ms -- every
    // pass re-reads the key array, so two halves is the sweet spot)
    int log2_passes = cap * sizeof(Slot) > (256ull << 20) ? 1 : 0;
    if (const char* e = getenv("DYD_DEDUP_PASSES_LOG2")) log2_passes = atoi(e);
    const int pass_shift = log2u(cap) - log2_passes;
    // the gated fallback normally exits at once: a small grid-stride grid keeps that exit cheap
    const unsigned grid = gate == nullptr ? grid_for(n) : std::min(grid_for(n), (unsigned)(NUM_SMS * 8));
    for (unsigned pass = 0; pass < (1u << log2_passes); ++pass) {
        dedup_insert_kernel<KIND><<<grid, HT_THREADS, url_pad(), s>>>(
            reinterpret_cast<const unsigned long long*>(d_keys), d_null, d_row_id, n, keep_mode, hdr, tab, cnt, shift, cap - 1, pass_shift, pass,
            gate, gate == nullptr);
        if (int rc = launch_check("dedup_insert_kernel")) return rc;
    }
    for (unsigned pass = 0; pass < (1u << log2_passes); ++pass) {
        dedup_lookup_kernel<KIND><<<grid, HT_THREADS, url_pad(), s>>>(
            reinterpret_cast<const unsigned long long*>(d_keys), d_null, d_row_id, n, keep_mode, hdr, tab, cnt, shift, cap - 1, pass_shift, pass, d_keep, d_rep,
            gate);
        if (int rc = launch_check("dedup_lookup_kernel")) return rc;
    }
    return 0;
}

static inline size_t dedup_table_bytes(int64_t n) {
    const uint64_t cap = table_capacity(n);
    return sizeof(TableHeader) + cap * sizeof(Slot) + cap * sizeof(unsigned);
}

template <int KIND>
static int dedup_impl(const uint64_t* d_keys, const uint8_t* d_null, const int64_t* d_row_id, int64_t n, int keep_mode,
                      uint8_t* d_keep, int64_t* d_rep, void* ws, size_t ws_bytes, void* stream) {
    DYD_REQUIRE(n >= 0 && n < (1LL << 40), DYD_E_ARG, "bad row count");
    DYD_REQUIRE(keep_mode >= 0 && keep_mode <= 2, DYD_E_ARG, "keep_mode must be 0 (first), 1 (last) or 2 (False)");
    if (n == 0) return 0;
    DYD_REQUIRE(d_keys && d_keep && d_rep && ws && (KIND != 1 || d_row_id), DYD_E_ARG, "null pointer");
    DYD_REQUIRE(((uintptr_t)ws & 15) == 0, DYD_E_ALIGN, "workspace must be 16-byte aligned");
    DYD_REQUIRE(ws_bytes >= dyd_dedup_workspace_bytes(n), DYD_E_WORKSPACE, "workspace too small");
    cudaStream_t s = as_stream(stream);
    if (!use_partitions(n)) return dedup_table<KIND>(d_keys, d_null, d_row_id, n, keep_mode, d_keep, d_rep, ws, nullptr, nullptr, s);

    const PartLayout L = part_layout(n, KIND != 0);
    char* base = reinterpret_cast<char*>(ws);
    TableHeader* hdr = reinterpret_cast<TableHeader*>(base);
    unsigned* cursors = reinterpret_cast<unsigned*>(base + L.cursors);
    int* overflow = reinterpret_cast<int*>(base + L.flag);
    ulonglong2* kv = reinterpret_cast<ulonglong2*>(base + L.kv);
    unsigned* rr = reinterpret_cast<unsigned*>(base + L.rr);
    const unsigned long long* k64 = reinterpret_cast<const unsigned long long*>(d_keys);
    DYD_CUDA(cudaMemsetAsync(base, 0xFF, sizeof(TableHeader), s));
    DYD_CUDA(cudaMemsetAsync(&hdr->null_count, 0, sizeof(unsigned long long), s));
    DYD_CUDA(cudaMemsetAsync(cursors, 0, L.kv - L.cursors, s));                       // cursors + overflow flag
    const int pshift = 64 - L.log2_np;
    dedup_partition_kernel<KIND><<<grid_for(n), HT_THREADS, url_pad(), s>>>(k64, d_null, d_row_id, n, hdr, cursors, overflow, kv, rr, pshift, L.pcap,
                                                                     d_keep, d_rep);
    if (int rc = launch_check("dedup_partition_kernel")) return rc;
    {
        const unsigned np = 1u << L.log2_np;
        constexpr bool R32 = KIND == 0;                // row ids are 0 .. n-1 < 2^31 on this path
        if (keep_mode == 0) dedup_resolve_kernel<KIND, 0, R32><<<np, PT_THREADS, url_pad(), s>>>(cursors, overflow, kv, rr, pshift, L.pcap, d_keep, d_rep);
        else if (keep_mode == 1) dedup_resolve_kernel<KIND, 1, R32><<<np, PT_THREADS, url_pad(), s>>>(cursors, overflow, kv, rr, pshift, L.pcap, d_keep, d_rep);
        else dedup_resolve_kernel<KIND, 2, R32><<<np, PT_THREADS, url_pad(), s>>>(cursors, overflow, kv, rr, pshift, L.pcap, d_keep, d_rep);
        if (int rc = launch_check("dedup_resolve_kernel")) return rc;
    }
    if (KIND == 0 && d_null != nullptr) {               // null cells: their answer needs the finished null statistics
        dedup_leftover_kernel<KIND><<<grid_for(n), HT_THREADS, url_pad(), s>>>(k64, d_null, d_row_id, n, keep_mode, hdr, overflow, d_keep, d_rep);
        if (int rc = launch_check("dedup_leftover_kernel")) return rc;
    }
    // a partition overflowed (one key repeated hundreds of times): the same call falls back on the device
    return dedup_table<KIND>(d_keys, d_null, d_row_id, n, keep_mode, d_keep, d_rep, base + L.fallback, overflow, hdr, s);
}


// ------------------------------------------------------------------------------- K4 + K5 joint, partitioned
// Steps 2 and 3 of the pipeline ask two questions about the same `source` key of every main row: which row represents its
// group (dedup, processor.py:140-144) and does the reference set hold it (anti-join, :194-199).  Both tables are scattered
// into the same key partitions; one CTA per partition then builds ONE shared-memory table holding, per distinct key, the
// first / last main row and the smallest reference row, and answers both questions for its main records.  Compared with
// the partitioned dedup followed by the global-table anti-join this drops the 128 MB table (fill, 5 M two-atomic inserts,
// 10 M random probes): 0.44 + 0.48 ms -> see DESIGN.md §4.
template <int KIND>
__global__ void __launch_bounds__(HT_THREADS)
ref_partition_kernel(const unsigned long long* __restrict__ keys, const uint8_t* __restrict__ null, int64_t n, unsigned* cursors,
                     int* overflow, ulonglong2* __restrict__ kv, int pshift, unsigned pcap) {
    const int64_t r = blockIdx.x * (int64_t)HT_THREADS + threadIdx.x;
    if (r >= n) return;
    unsigned long long key, id;
    if (KIND == 0) {
        if (null != nullptr && null[r] != 0) return;                 // ref.dropna()
        key = norm_key(keys[r]); id = (unsigned long long)r;
    } else {
        const ulonglong2 rec = reinterpret_cast<const ulonglong2*>(keys)[r];
        if ((long long)rec.y < 0) return;
        key = norm_key(rec.x); id = rec.y;
    }
    const unsigned p = pshift >= 64 ? 0u : (unsigned)((key * GOLD) >> pshift);
    const unsigned slot = atomicAdd(&cursors[p], 1u);
    if (slot >= pcap) { *overflow = 1; return; }
    kv[(size_t)p * pcap + slot] = make_ulonglong2(key, id);
}

// (key, id) records whose id is set back to padding once the step has read them (peer-memory exchange buffers)
__global__ void __launch_bounds__(HT_THREADS)
reset_records_kernel(unsigned long long* __restrict__ records, int64_t m) {
    const int64_t r = blockIdx.x * (int64_t)HT_THREADS + threadIdx.x;
    if (r < m) records[2 * r + 1] = ~0ULL;
}

template <int KIND, int MODE, bool ROW32>
__global__ void __launch_bounds__(PT_THREADS)
joint_resolve_kernel(const unsigned* __restrict__ cur_m, const unsigned* __restrict__ cur_r, const int* __restrict__ overflow,
                     const ulonglong2* __restrict__ kv_m, const unsigned* __restrict__ rr_m, const ulonglong2* __restrict__ kv_r,
                     int pshift, unsigned pcap_m, unsigned pcap_r,
                     uint8_t* __restrict__ keep, int64_t* __restrict__ rep, uint8_t* __restrict__ keep2, int64_t* __restrict__ ref2) {
    using RowT = typename std::conditional<ROW32, unsigned, unsigned long long>::type;
    using SRowT = typename std::conditional<ROW32, int, long long>::type;
    __shared__ unsigned long long skey[PT_SLOTS];
    __shared__ RowT srow[PT_SLOTS];                    // first (min) or last (max) main row of the key
    __shared__ RowT sref[PT_SLOTS];                    // smallest reference row of the key, all ones = the reference set lacks it
    __shared__ unsigned scnt[MODE == 2 ? PT_SLOTS : 1];
    if (*overflow) return;                             // the gated global-table kernels take over
    const unsigned p = blockIdx.x;
    const unsigned cm = min(cur_m[p], pcap_m), cr = min(cur_r[p], pcap_r);
    if (cm == 0) return;                               // no main record asks anything here
    for (int i = threadIdx.x; i < PT_SLOTS; i += PT_THREADS) {
        skey[i] = EMPTY; srow[i] = (RowT)~(RowT)0; sref[i] = (RowT)~(RowT)0;
        if (MODE == 2) scnt[i] = 0;
    }
    __syncthreads();
    const ulonglong2* refs = kv_r + (size_t)p * pcap_r;
    for (unsigned i = threadIdx.x; i < cr; i += PT_THREADS) {
        const ulonglong2 rec = refs[i];
        unsigned s = (unsigned)(((rec.x * GOLD) << (64 - pshift)) >> 54) & (PT_SLOTS - 1);
        for (;;) {
            const unsigned long long prev = atomicCAS(&skey[s], EMPTY, rec.x);
            if (prev == EMPTY || prev == rec.x) break;
            s = (s + 1) & (PT_SLOTS - 1);
        }
        atomicMin(&sref[s], (RowT)rec.y);
    }
    const ulonglong2* mine = kv_m + (size_t)p * pcap_m;
    unsigned long long id[PT_MAX_PER_THREAD];
    int at[PT_MAX_PER_THREAD];
#pragma unroll
    for (int u = 0; u < PT_MAX_PER_THREAD; ++u) {
        const unsigned i = threadIdx.x + u * PT_THREADS;
        at[u] = -1;
        if (i < cm) {
            const ulonglong2 rec = mine[i];
            id[u] = rec.y;
            unsigned s = (unsigned)(((rec.x * GOLD) << (64 - pshift)) >> 54) & (PT_SLOTS - 1);
            for (;;) {
                const unsigned long long prev = atomicCAS(&skey[s], EMPTY, rec.x);
                if (prev == EMPTY || prev == rec.x) break;
                s = (s + 1) & (PT_SLOTS - 1);
            }
            if (MODE == 1) atomicMax(reinterpret_cast<SRowT*>(&srow[s]), (SRowT)rec.y);
            else atomicMin(&srow[s], (RowT)rec.y);
            if (MODE == 2) atomicAdd(&scnt[s], 1u);
            at[u] = (int)s;
        }
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < PT_MAX_PER_THREAD; ++u) {
        if (at[u] < 0) continue;
        const unsigned i = threadIdx.x + u * PT_THREADS;
        const long long rp = (long long)(SRowT)srow[at[u]];
        const bool kept = MODE == 2 ? (scnt[at[u]] == 1u) : (rp == (long long)id[u]);
        const RowT hit = sref[at[u]];
        const bool in_ref = hit != (RowT)~(RowT)0;
        if (kept && rp == (long long)id[u] && !in_ref) continue;       // both answers were written by the partition kernel
        const size_t out = KIND == 0 ? (size_t)id[u] : (size_t)rr_m[(size_t)p * pcap_m + i];
        if (!(kept && rp == (long long)id[u])) { rep[out] = rp; keep[out] = kept ? 1 : 0; }
        if (in_ref) { keep2[out] = 0; ref2[out] = (long long)hit; }
    }
}

}  // namespace dyd

using namespace dyd;

extern "C" int dyd_hash_strings(const int64_t* d_off, const uint8_t* d_bytes, int64_t n, uint64_t* d_hash, void* stream) {
    DYD_REQUIRE(n >= 0, DYD_E_ARG, "negative count");
    if (n == 0) return 0;
    DYD_REQUIRE(d_off && d_hash, DYD_E_ARG, "null pointer");
    hash_strings_kernel<<<grid_for(n), HT_THREADS, url_pad(), as_stream(stream)>>>(d_off, d_bytes, n, d_hash);
    return launch_check("hash_strings_kernel");
}

extern "C" size_t dyd_dedup_workspace_bytes(int64_t n) {
    const size_t table = dedup_table_bytes(n);
    return use_partitions(n) ? part_layout(n, true).total : table;
}

extern "C" int dyd_dedup(const uint64_t* d_keys, const uint8_t* d_null, int64_t n, int keep_mode,
                         uint8_t* d_keep, int64_t* d_rep, void* ws, size_t ws_bytes, void* stream) {
    return dedup_impl<0>(d_keys, d_null, nullptr, n, keep_mode, d_keep, d_rep, ws, ws_bytes, stream);
}

extern "C" int dyd_dedup_ids(const uint64_t* d_keys, const int64_t* d_row_id, int64_t n, int keep_mode,
                             uint8_t* d_keep, int64_t* d_rep, void* ws, size_t ws_bytes, void* stream) {
    return dedup_impl<1>(d_keys, nullptr, d_row_id, n, keep_mode, d_keep, d_rep, ws, ws_bytes, stream);
}

extern "C" int dyd_dedup_records(const int64_t* d_records, int64_t m, int keep_mode,
                                 uint8_t* d_keep, int64_t* d_rep, void* ws, size_t ws_bytes, void* stream) {
    return dedup_impl<2>(reinterpret_cast<const uint64_t*>(d_records), nullptr, nullptr, m, keep_mode, d_keep, d_rep, ws, ws_bytes, stream);
}

extern "C" int dyd_shard_bucket(const uint64_t* d_keys, const uint8_t* d_null, int64_t row_base, int64_t n, int32_t world,
                                int64_t cap, int64_t* d_records, uint64_t* d_cursors, int32_t* d_overflow, void* stream) {
    DYD_REQUIRE(n >= 0 && world >= 1 && cap >= 0, DYD_E_ARG, "bad arguments");
    DYD_REQUIRE(d_records && d_cursors && d_overflow && (n == 0 || d_keys), DYD_E_ARG, "null pointer");
    cudaStream_t s = as_stream(stream);
    DYD_CUDA(cudaMemsetAsync(d_records, 0xFF, sizeof(int64_t) * 2 * (size_t)world * (size_t)cap, s));
    DYD_CUDA(cudaMemsetAsync(d_cursors, 0, sizeof(uint64_t) * world, s));
    DYD_CUDA(cudaMemsetAsync(d_overflow, 0, sizeof(int32_t), s));
    if (n == 0) return 0;
    shard_bucket_kernel<<<grid_for(n), HT_THREADS, url_pad(), s>>>(reinterpret_cast<const unsigned long long*>(d_keys), d_null, row_base, n,
                                                          world, cap, reinterpret_cast<long long*>(d_records),
                                                          reinterpret_cast<unsigned long long*>(d_cursors), d_overflow);
    return launch_check("shard_bucket_kernel");
}

static int shard_bucket_p2p_impl(const uint64_t* d_keys, const uint8_t* d_null, int64_t row_base, int64_t n, int32_t world,
                                 int32_t my_rank, int64_t cap, int64_t* const* d_peer_records, uint32_t* d_sent_row,
                                 uint64_t* d_cursors, int32_t* d_overflow, uint8_t* d_def_keep_d, int64_t* d_def_rep_d,
                                 uint8_t* d_def_keep_a, int64_t* d_def_ref_row, void* stream);
extern "C" int dyd_shard_bucket_p2p(const uint64_t* d_keys, const uint8_t* d_null, int64_t row_base, int64_t n, int32_t world,
                                    int32_t my_rank, int64_t cap, int64_t* const* d_peer_records, uint32_t* d_sent_row,
                                    uint64_t* d_cursors, int32_t* d_overflow, void* stream) {
    return shard_bucket_p2p_impl(d_keys, d_null, row_base, n, world, my_rank, cap, d_peer_records, d_sent_row, d_cursors, d_overflow,
                                 nullptr, nullptr, nullptr, nullptr, stream);
}
extern "C" int dyd_shard_bucket_p2p_defaults(const uint64_t* d_keys, const uint8_t* d_null, int64_t row_base, int64_t n, int32_t world,
                                             int32_t my_rank, int64_t cap, int64_t* const* d_peer_records, uint32_t* d_sent_row,
                                             uint64_t* d_cursors, int32_t* d_overflow, uint8_t* d_keep_dedup, int64_t* d_rep_dedup,
                                             uint8_t* d_keep_anti, int64_t* d_ref_row, void* stream) {
    DYD_REQUIRE(world <= P2P_MAX_WORLD, DYD_E_ARG, "the sparse-reply form needs world <= 64");
    DYD_REQUIRE(d_keep_dedup && d_rep_dedup && d_keep_anti && d_ref_row, DYD_E_ARG, "null pointer");
    return shard_bucket_p2p_impl(d_keys, d_null, row_base, n, world, my_rank, cap, d_peer_records, d_sent_row, d_cursors, d_overflow,
                                 d_keep_dedup, d_rep_dedup, d_keep_anti, d_ref_row, stream);
}
static int shard_bucket_p2p_impl(const uint64_t* d_keys, const uint8_t* d_null, int64_t row_base, int64_t n, int32_t world,
                                 int32_t my_rank, int64_t cap, int64_t* const* d_peer_records, uint32_t* d_sent_row,
                                 uint64_t* d_cursors, int32_t* d_overflow, uint8_t* d_def_keep_d, int64_t* d_def_rep_d,
                                 uint8_t* d_def_keep_a, int64_t* d_def_ref_row, void* stream) {
    DYD_REQUIRE(n >= 0 && n < (1LL << 32) && world >= 1 && cap >= 0 && my_rank >= 0 && my_rank < world, DYD_E_ARG, "bad arguments");
    DYD_REQUIRE(d_peer_records && d_cursors && d_overflow && (n == 0 || d_keys), DYD_E_ARG, "null pointer");
    cudaStream_t s = as_stream(stream);
    DYD_CUDA(cudaMemsetAsync(d_cursors, 0, sizeof(uint64_t) * world, s));
    DYD_CUDA(cudaMemsetAsync(d_overflow, 0, sizeof(int32_t), s));
    if (n == 0) return 0;
    const char* e = getenv("DYD_SCATTER_STAGED");
    if (world <= P2P_MAX_WORLD && (d_def_keep_d != nullptr || !(e && atoi(e) == 0))) {
        shard_bucket_staged_kernel<<<(unsigned)((n + ST_PER_BLOCK - 1) / ST_PER_BLOCK), ST_THREADS, 0, s>>>(
            reinterpret_cast<const unsigned long long*>(d_keys), d_null, row_base, n, world, my_rank, cap,
            reinterpret_cast<long long* const*>(d_peer_records), d_sent_row, reinterpret_cast<unsigned long long*>(d_cursors), d_overflow,
            d_def_keep_d, d_def_rep_d, d_def_keep_a, d_def_ref_row);
        return launch_check("shard_bucket_staged_kernel");
    }
    shard_bucket_p2p_kernel<<<grid_for(n), HT_THREADS, url_pad(), s>>>(reinterpret_cast<const unsigned long long*>(d_keys), d_null, row_base, n,
                                                              world, my_rank, cap, reinterpret_cast<long long* const*>(d_peer_records),
                                                              d_sent_row, reinterpret_cast<unsigned long long*>(d_cursors), d_overflow);
    return launch_check("shard_bucket_p2p_kernel");
}

extern "C" int dyd_shard_pack_reply_p2p(int64_t* d_records, const uint8_t* d_keep, const int64_t* d_rep, int64_t m, int64_t cap,
                                        int32_t my_rank, int64_t* const* d_peer_reply, int32_t mode, int32_t reset_records, void* stream) {
    DYD_REQUIRE(m >= 0 && cap > 0 && my_rank >= 0 && m % cap == 0 && (mode == 0 || mode == 1), DYD_E_ARG, "bad arguments");
    if (m == 0) return 0;
    DYD_REQUIRE(d_records && d_keep && d_rep && d_peer_reply, DYD_E_ARG, "null pointer");
    shard_pack_reply_p2p_kernel<<<grid_for(m), HT_THREADS, url_pad(), as_stream(stream)>>>(reinterpret_cast<long long*>(d_records), d_keep, d_rep,
                                                                                    m, cap, my_rank, reinterpret_cast<long long* const*>(d_peer_reply),
                                                                                    mode, reset_records != 0);
    return launch_check("shard_pack_reply_p2p_kernel");
}

extern "C" int dyd_shard_unpack_p2p(const int64_t* d_reply, const uint32_t* d_sent_row, const uint64_t* d_cursors, int32_t world,
                                    int64_t cap, int64_t n, uint8_t* d_keep, int64_t* d_rep, int32_t mode, void* stream) {
    DYD_REQUIRE(world >= 1 && cap >= 0 && n >= 0 && (mode == 0 || mode == 1), DYD_E_ARG, "bad arguments");
    if (cap == 0 || n == 0) return 0;
    DYD_REQUIRE(d_reply && d_sent_row && d_cursors && d_keep && d_rep, DYD_E_ARG, "null pointer");
    shard_unpack_p2p_kernel<<<grid_for((int64_t)world * cap), HT_THREADS, url_pad(), as_stream(stream)>>>(
        reinterpret_cast<const long long*>(d_reply), d_sent_row, reinterpret_cast<const unsigned long long*>(d_cursors), world, cap, n, d_keep, d_rep, mode);
    return launch_check("shard_unpack_p2p_kernel");
}

extern "C" int dyd_shard_pack_reply2_p2p(int64_t* d_records, const uint8_t* d_keep_dedup, const int64_t* d_rep_dedup,
                                         const uint8_t* d_keep_anti, const int64_t* d_ref_row, int64_t m, int64_t cap, int32_t my_rank,
                                         int64_t* const* d_peer_reply2, int32_t reset_records, int32_t sparse, void* stream) {
    DYD_REQUIRE(m >= 0 && cap > 0 && my_rank >= 0 && m % cap == 0, DYD_E_ARG, "bad arguments");
    if (m == 0) return 0;
    DYD_REQUIRE(d_records && d_keep_dedup && d_rep_dedup && d_keep_anti && d_ref_row && d_peer_reply2, DYD_E_ARG, "null pointer");
    shard_pack_reply2_p2p_kernel<<<grid_for(m), HT_THREADS, 0, as_stream(stream)>>>(reinterpret_cast<long long*>(d_records), d_keep_dedup, d_rep_dedup,
                                                                                   d_keep_anti, d_ref_row, m, cap, my_rank,
                                                                                   reinterpret_cast<long long* const*>(d_peer_reply2), reset_records != 0, sparse != 0);
    return launch_check("shard_pack_reply2_p2p_kernel");
}

extern "C" int dyd_shard_unpack2_p2p(int64_t* d_reply2, const uint32_t* d_sent_row, const uint64_t* d_cursors, int32_t world,
                                     int64_t cap, int64_t n, uint8_t* d_keep_dedup, int64_t* d_rep_dedup, uint8_t* d_keep_anti,
                                     int64_t* d_ref_row, int32_t sparse, void* stream) {
    DYD_REQUIRE(world >= 1 && cap >= 0 && n >= 0, DYD_E_ARG, "bad arguments");
    if (cap == 0 || n == 0) return 0;
    DYD_REQUIRE(d_reply2 && d_sent_row && d_cursors && d_keep_dedup && d_rep_dedup && d_keep_anti && d_ref_row, DYD_E_ARG, "null pointer");
    DYD_REQUIRE(((uintptr_t)d_reply2 & 15) == 0, DYD_E_ALIGN, "reply buffer must be 16-byte aligned");
    shard_unpack2_p2p_kernel<<<grid_for((int64_t)world * cap), HT_THREADS, 0, as_stream(stream)>>>(
        reinterpret_cast<longlong2*>(d_reply2), d_sent_row, reinterpret_cast<const unsigned long long*>(d_cursors), world, cap, n,
        d_keep_dedup, d_rep_dedup, d_keep_anti, d_ref_row, sparse != 0);
    return launch_check("shard_unpack2_p2p_kernel");
}

extern "C" int dyd_shard_pack_reply(const int64_t* d_records, const uint8_t* d_keep, const int64_t* d_rep, int64_t m,
                                    int64_t* d_reply, int32_t mode, void* stream) {
    DYD_REQUIRE(m >= 0 && (mode == 0 || mode == 1), DYD_E_ARG, "bad arguments");
    if (m == 0) return 0;
    DYD_REQUIRE(d_records && d_keep && d_rep && d_reply, DYD_E_ARG, "null pointer");
    shard_pack_reply_kernel<<<grid_for(m), HT_THREADS, url_pad(), as_stream(stream)>>>(reinterpret_cast<const long long*>(d_records), d_keep, d_rep, m,
                                                                                reinterpret_cast<long long*>(d_reply), mode);
    return launch_check("shard_pack_reply_kernel");
}

extern "C" int dyd_shard_unpack(const int64_t* d_reply, int64_t m, int64_t row_base, int64_t n,
                                uint8_t* d_keep, int64_t* d_rep, int32_t mode, void* stream) {
    DYD_REQUIRE(m >= 0 && n >= 0 && (mode == 0 || mode == 1), DYD_E_ARG, "bad arguments");
    if (m == 0) return 0;
    DYD_REQUIRE(d_reply && d_keep && d_rep, DYD_E_ARG, "null pointer");
    shard_unpack_kernel<<<grid_for(m), HT_THREADS, url_pad(), as_stream(stream)>>>(reinterpret_cast<const long long*>(d_reply), m, row_base, n, d_keep, d_rep, mode);
    return launch_check("shard_unpack_kernel");
}

static inline uint64_t antijoin_capacity(int64_t n);
extern "C" size_t dyd_antijoin_workspace_bytes(int64_t n_ref) {
    return antijoin_capacity(n_ref < 0 ? 0 : n_ref) * sizeof(Slot);
}

// The anti-join table is sized for a load factor <= 2/3 (the dedup table, whose every lookup is a hit after an insert, for
// <= 1/2): for 5 M reference keys that is 128 MB instead of 256 MB to clear, build and probe.
static inline uint64_t antijoin_capacity(int64_t n) {
    uint64_t cap = 1024;
    while (cap * 2 < (uint64_t)(n > 0 ? n : 0) * 3) cap <<= 1;
    return cap;
}
static inline size_t antijoin_table_bytes(int64_t n_ref) { return antijoin_capacity(n_ref) * sizeof(Slot); }

// the global-table anti-join; gate == nullptr: unconditional, else only if *gate != 0
template <int KIND>
static int antijoin_table(uint64_t* d_ref, const uint8_t* d_ref_null, int64_t n_ref, const uint64_t* d_main, const uint8_t* d_main_null,
                          int64_t n_main, uint8_t* d_keep, int64_t* d_ref_row, void* ws, bool reset_ref, const int* gate, cudaStream_t s) {
    const uint64_t cap = antijoin_capacity(n_ref);
    const int shift = 64 - log2u(cap);
    Slot* tab = reinterpret_cast<Slot*>(ws);
    if (gate == nullptr) DYD_CUDA(cudaMemsetAsync(ws, 0xFF, cap * sizeof(Slot), s));
    else {
        gated_fill_kernel<<<NUM_SMS * 8, 256, url_pad(), s>>>(reinterpret_cast<uint4*>(tab), cap * sizeof(Slot) / 16, 0xFFFFFFFFu, gate);
        if (int rc = launch_check("gated_fill_kernel")) return rc;
    }
    if (n_ref > 0) {
        antijoin_build_kernel<KIND><<<grid_for(n_ref), HT_THREADS, url_pad(), s>>>(
            reinterpret_cast<unsigned long long*>(d_ref), d_ref_null, n_ref, tab, shift, cap - 1, reset_ref, gate);
        if (int rc = launch_check("antijoin_build_kernel")) return rc;
    }
    if (n_main == 0) return 0;
    antijoin_probe_kernel<KIND><<<grid_for(n_main), HT_THREADS, url_pad(), s>>>(
        reinterpret_cast<const unsigned long long*>(d_main), d_main_null, n_main, tab, shift, cap - 1, d_keep, d_ref_row, gate);
    return launch_check("antijoin_probe_kernel");
}

template <int KIND>
static int antijoin_impl(uint64_t* d_ref, const uint8_t* d_ref_null, int64_t n_ref, const uint64_t* d_main, const uint8_t* d_main_null,
                         int64_t n_main, uint8_t* d_keep, int64_t* d_ref_row, void* ws, size_t ws_bytes, bool reset_ref, void* stream) {
    DYD_REQUIRE(n_main >= 0 && n_ref >= 0, DYD_E_ARG, "negative count");
    if (n_main == 0 && !(KIND == 2 && reset_ref)) return 0;
    DYD_REQUIRE((n_main == 0 || (d_main && d_keep && d_ref_row)) && ws && (n_ref == 0 || d_ref), DYD_E_ARG, "null pointer");
    DYD_REQUIRE(((uintptr_t)ws & 15) == 0, DYD_E_ALIGN, "workspace must be 16-byte aligned");
    DYD_REQUIRE(KIND == 0 || ((((uintptr_t)d_ref | (uintptr_t)d_main) & 15) == 0), DYD_E_ALIGN, "records must be 16-byte aligned");
    DYD_REQUIRE(ws_bytes >= dyd_antijoin_workspace_bytes(n_ref), DYD_E_WORKSPACE, "workspace too small");
    cudaStream_t s = as_stream(stream);
    return antijoin_table<KIND>(d_ref, d_ref_null, n_ref, d_main, d_main_null, n_main, d_keep, d_ref_row, ws, reset_ref, nullptr, s);
}

extern "C" int dyd_antijoin(const uint64_t* d_main_keys, const uint8_t* d_main_null, int64_t n_main,
                            const uint64_t* d_ref_keys, const uint8_t* d_ref_null, int64_t n_ref,
                            uint8_t* d_keep, int64_t* d_ref_row, void* ws, size_t ws_bytes, void* stream) {
    return antijoin_impl<0>(const_cast<uint64_t*>(d_ref_keys), d_ref_null, n_ref, d_main_keys, d_main_null, n_main, d_keep, d_ref_row,
                            ws, ws_bytes, false, stream);
}

extern "C" int dyd_antijoin_records(int64_t* d_ref_records, int64_t m_ref, const int64_t* d_main_records, int64_t m_main,
                                    uint8_t* d_keep, int64_t* d_ref_row, void* ws, size_t ws_bytes, int32_t reset_ref, void* stream) {
    return antijoin_impl<2>(reinterpret_cast<uint64_t*>(d_ref_records), nullptr, m_ref, reinterpret_cast<const uint64_t*>(d_main_records), nullptr,
                            m_main, d_keep, d_ref_row, ws, ws_bytes, reset_ref != 0, stream);
}


// ------------------------------------------------------------------------------- K4 + K5 joint: host side
extern "C" size_t dyd_url_filter_workspace_bytes(int64_t n_main, int64_t n_ref);
#ifndef DYD_JOINT_FILL
#define DYD_JOINT_FILL 192
#endif
struct JointLayout {
    int log2_np;
    unsigned pcap_m, pcap_r;
    size_t cur_m, cur_r, flag, kv_m, rr_m, kv_r, fb_dedup, fb_anti, total;
};
static inline JointLayout joint_layout(int64_t n_main, int64_t n_ref, bool with_r) {
    JointLayout L{};
    int k = 0;
    while ((((long long)DYD_JOINT_FILL << k) < n_main || (((long long)DYD_JOINT_FILL / 2) << k) < n_ref) && k < 30) ++k;   // main fill < 192, reference fill < 96 on average
    L.log2_np = k;
    const uint64_t np = 1ULL << k;
    L.pcap_m = (unsigned)std::min<uint64_t>(PT_THREADS * PT_MAX_PER_THREAD - 1, 2 * ((uint64_t)n_main / np) + 32);
    L.pcap_r = (unsigned)std::min<uint64_t>(383, 2 * ((uint64_t)n_ref / np) + 32);      // pcap_m + pcap_r < PT_SLOTS
    size_t o = sizeof(TableHeader);
    L.cur_m = o; o += sizeof(unsigned) * np;
    L.cur_r = o; o += sizeof(unsigned) * np;
    L.flag = o; o += 16;
    o = (o + 15) & ~(size_t)15;
    L.kv_m = o; o += sizeof(ulonglong2) * np * L.pcap_m;
    L.kv_r = o; o += sizeof(ulonglong2) * np * L.pcap_r;
    L.rr_m = o; o += with_r ? sizeof(unsigned) * np * L.pcap_m : 0;
    o = (o + 255) & ~(size_t)255;
    L.fb_dedup = o; o += dedup_table_bytes(n_main);
    o = (o + 255) & ~(size_t)255;
    L.fb_anti = o; o += antijoin_table_bytes(n_ref);
    L.total = o;
    return L;
}

// dedup + anti-join of the same main keys; KIND 0: key arrays with null flags, KIND 2: (key, id) records of the exchange
template <int KIND>
static int url_filter_impl(const uint64_t* d_main, const uint8_t* d_main_null, int64_t n_main, uint64_t* d_ref, const uint8_t* d_ref_null,
                           int64_t n_ref, int keep_mode, uint8_t* d_keep, int64_t* d_rep, uint8_t* d_keep2, int64_t* d_ref2,
                           void* ws, size_t ws_bytes, bool reset_ref, int64_t id_bound, void* stream) {
    DYD_REQUIRE(n_main >= 0 && n_ref >= 0 && n_main < (1LL << 40) && n_ref < (1LL << 40), DYD_E_ARG, "bad row count");
    DYD_REQUIRE(keep_mode >= 0 && keep_mode <= 2, DYD_E_ARG, "keep_mode must be 0 (first), 1 (last) or 2 (False)");
    cudaStream_t s = as_stream(stream);
    if (n_main > 0) {
        DYD_REQUIRE(d_main && d_keep && d_rep && d_keep2 && d_ref2 && ws && (n_ref == 0 || d_ref), DYD_E_ARG, "null pointer");
        DYD_REQUIRE(((uintptr_t)ws & 15) == 0, DYD_E_ALIGN, "workspace must be 16-byte aligned");
        DYD_REQUIRE(ws_bytes >= dyd_url_filter_workspace_bytes(n_main, n_ref), DYD_E_WORKSPACE, "workspace too small");
        char* base = reinterpret_cast<char*>(ws);
        if (!use_partitions(n_main) || n_ref >= (1LL << 32) - 1) {
            // small tables: the two single-question kernels one after the other (their tables sit in L2)
            if (int rc = dedup_impl<KIND>(d_main, d_main_null, nullptr, n_main, keep_mode, d_keep, d_rep, ws, dyd_dedup_workspace_bytes(n_main), stream)) return rc;
            char* aws = base + ((dyd_dedup_workspace_bytes(n_main) + 255) & ~(size_t)255);
            if (int rc = antijoin_table<KIND>(d_ref, d_ref_null, n_ref, d_main, d_main_null, n_main, d_keep2, d_ref2, aws, false, nullptr, s)) return rc;
        } else {
            const JointLayout L = joint_layout(n_main, n_ref, KIND != 0);
            TableHeader* hdr = reinterpret_cast<TableHeader*>(base);
            unsigned* cur_m = reinterpret_cast<unsigned*>(base + L.cur_m);
            unsigned* cur_r = reinterpret_cast<unsigned*>(base + L.cur_r);
            int* overflow = reinterpret_cast<int*>(base + L.flag);
            ulonglong2* kv_m = reinterpret_cast<ulonglong2*>(base + L.kv_m);
            ulonglong2* kv_r = reinterpret_cast<ulonglong2*>(base + L.kv_r);
            unsigned* rr_m = reinterpret_cast<unsigned*>(base + L.rr_m);
            const unsigned long long* m64 = reinterpret_cast<const unsigned long long*>(d_main);
            DYD_CUDA(cudaMemsetAsync(base, 0xFF, sizeof(TableHeader), s));
            DYD_CUDA(cudaMemsetAsync(&hdr->null_count, 0, sizeof(unsigned long long), s));
            DYD_CUDA(cudaMemsetAsync(cur_m, 0, L.kv_m - L.cur_m, s));                  // both cursor arrays + overflow flag
            const int pshift = 64 - L.log2_np;
            if (n_ref > 0) {
                ref_partition_kernel<KIND><<<grid_for(n_ref), HT_THREADS, 0, s>>>(reinterpret_cast<const unsigned long long*>(d_ref), d_ref_null, n_ref,
                                                                                  cur_r, overflow, kv_r, pshift, L.pcap_r);
                if (int rc = launch_check("ref_partition_kernel")) return rc;
            }
            dedup_partition_kernel<KIND><<<grid_for(n_main), HT_THREADS, 0, s>>>(m64, d_main_null, nullptr, n_main, hdr, cur_m, overflow, kv_m, rr_m, pshift,
                                                                               L.pcap_m, d_keep, d_rep, d_keep2, d_ref2);
            if (int rc = launch_check("dedup_partition_kernel")) return rc;
            const unsigned np = 1u << L.log2_np;
            // 32-bit rows in the shared-memory tables (native shared atomics) when every id fits: always for key arrays (main rows
            // < 2^31, reference rows < 2^32 - 1 on this path), for exchange records when the caller bounds the ids below 2^31
            const bool r32 = KIND == 0 || (id_bound > 0 && id_bound < (1LL << 31));
#define DYD_JOINT(MODE, R32) joint_resolve_kernel<KIND, MODE, R32><<<np, PT_THREADS, 0, s>>>(cur_m, cur_r, overflow, kv_m, rr_m, kv_r, pshift, L.pcap_m, \
                                                                                             L.pcap_r, d_keep, d_rep, d_keep2, d_ref2)
            if (r32) { if (keep_mode == 0) DYD_JOINT(0, true); else if (keep_mode == 1) DYD_JOINT(1, true); else DYD_JOINT(2, true); }
            else if (KIND != 0) { if (keep_mode == 0) DYD_JOINT(0, false); else if (keep_mode == 1) DYD_JOINT(1, false); else DYD_JOINT(2, false); }
#undef DYD_JOINT
            if (int rc = launch_check("joint_resolve_kernel")) return rc;
            if (KIND == 0 && d_main_null != nullptr) {
                dedup_leftover_kernel<KIND><<<grid_for(n_main), HT_THREADS, 0, s>>>(m64, d_main_null, nullptr, n_main, keep_mode, hdr, overflow, d_keep, d_rep);
                if (int rc = launch_check("dedup_leftover_kernel")) return rc;
            }
            // a partition overflowed (one key repeated hundreds of times): both questions are asked again of the global tables
            if (int rc = dedup_table<KIND>(d_main, d_main_null, nullptr, n_main, keep_mode, d_keep, d_rep, base + L.fb_dedup, overflow, hdr, s)) return rc;
            if (int rc = antijoin_table<KIND>(d_ref, d_ref_null, n_ref, d_main, d_main_null, n_main, d_keep2, d_ref2, base + L.fb_anti, false, overflow, s)) return rc;
        }
    }
    if (KIND == 2 && reset_ref && n_ref > 0) {
        reset_records_kernel<<<grid_for(n_ref), HT_THREADS, 0, s>>>(reinterpret_cast<unsigned long long*>(d_ref), n_ref);
        if (int rc = launch_check("reset_records_kernel")) return rc;
    }
    return 0;
}

extern "C" size_t dyd_url_filter_workspace_bytes(int64_t n_main, int64_t n_ref) {
    if (n_main < 0) n_main = 0;
    if (n_ref < 0) n_ref = 0;
    const size_t separate = ((dyd_dedup_workspace_bytes(n_main) + 255) & ~(size_t)255) + antijoin_table_bytes(n_ref);
    if (!use_partitions(n_main) || n_ref >= (1LL << 32) - 1) return separate;
    return std::max(separate, joint_layout(n_main, n_ref, true).total);
}

extern "C" int dyd_url_filter(const uint64_t* d_main_keys, const uint8_t* d_main_null, int64_t n_main,
                              const uint64_t* d_ref_keys, const uint8_t* d_ref_null, int64_t n_ref, int keep_mode,
                              uint8_t* d_keep, int64_t* d_rep, uint8_t* d_keep_ref, int64_t* d_ref_row,
                              void* ws, size_t ws_bytes, void* stream) {
    return url_filter_impl<0>(d_main_keys, d_main_null, n_main, const_cast<uint64_t*>(d_ref_keys), d_ref_null, n_ref, keep_mode, d_keep, d_rep,
                              d_keep_ref, d_ref_row, ws, ws_bytes, false, 0, stream);
}

extern "C" int dyd_url_filter_records(int64_t* d_ref_records, int64_t m_ref, const int64_t* d_main_records, int64_t m_main, int keep_mode,
                                      uint8_t* d_keep, int64_t* d_rep, uint8_t* d_keep_ref, int64_t* d_ref_row,
                                      void* ws, size_t ws_bytes, int32_t reset_ref, int64_t id_bound, void* stream) {
    return url_filter_impl<2>(reinterpret_cast<const uint64_t*>(d_main_records), nullptr, m_main, reinterpret_cast<uint64_t*>(d_ref_records), nullptr,
                              m_ref, keep_mode, d_keep, d_rep, d_keep_ref, d_ref_row, ws, ws_bytes, reset_ref != 0, id_bound, stream);
}

// K3 label remap through lookup tables, K6 label->category expansion (stable partition),
// split assignment from the host permutation, and the YOLO cx/cy/w/h arithmetic.
#include "common.cuh"

namespace dyd {

constexpr int LB_THREADS = 256;
constexpr int IMG_PER_CTA = 256;      // K3: images per block
constexpr int IMG_PER_WARP = 256;     // K6: images per warp chunk
constexpr int MAX_CAT = 256;

// ------------------------------------------------------------------------------- K3
__global__ void __launch_bounds__(LB_THREADS)
label_lut_kernel(const int64_t* __restrict__ img_off, const int32_t* __restrict__ label_id, int64_t n_img,
                 const int32_t* __restrict__ lut_new, const int32_t* __restrict__ lut_ntok,
                 const int32_t* __restrict__ lut_nrep, int32_t n_vocab,
                 int32_t* __restrict__ new_id, uint8_t* __restrict__ row_rep, unsigned long long* counters) {
    const int64_t i0 = blockIdx.x * (int64_t)IMG_PER_CTA;
    const int64_t i1 = min(i0 + IMG_PER_CTA, n_img);
    const int64_t q0 = img_off[i0], q1 = img_off[i1];
    unsigned long long c_obj = 0, c_miss = 0, c_lab = 0, c_rlab = 0, c_robj = 0, c_rrow = 0;
    auto remap = [&](int32_t v) {
        int32_t out = v;
        ++c_obj;
        if (v < 0 || v >= n_vocab) { ++c_miss; }
        else {
            c_lab += (unsigned)__ldg(lut_ntok + v);
            const int32_t nr = __ldg(lut_nrep + v);
            if (nr > 0) { out = __ldg(lut_new + v); c_rlab += (unsigned)nr; ++c_robj; }
        }
        return out;
    };
    // phase 1: coalesced sweep over the block's objects, four labels per thread and access where
    // the arrays allow 128-bit accesses (the ragged ends of the block's range go one by one)
    const bool vec = ((reinterpret_cast<uintptr_t>(label_id) | reinterpret_cast<uintptr_t>(new_id)) & 15) == 0;
    const int64_t qa = vec ? min(q1, (q0 + 3) & ~(int64_t)3) : q1, qb = vec ? qa + ((q1 - qa) & ~(int64_t)3) : q1;
    for (int64_t q = q0 + threadIdx.x; q < qa; q += LB_THREADS) new_id[q] = remap(label_id[q]);
    for (int64_t q = qa + 4 * (int64_t)threadIdx.x; q < qb; q += 4 * LB_THREADS) {
        const int4 v = *reinterpret_cast<const int4*>(label_id + q);
        *reinterpret_cast<int4*>(new_id + q) = make_int4(remap(v.x), remap(v.y), remap(v.z), remap(v.w));
    }
    for (int64_t q = qb + threadIdx.x; q < q1; q += LB_THREADS) new_id[q] = remap(label_id[q]);
    // phase 2: one thread per image decides row_replaced (labels are L1/L2 hot from phase 1)
    const int64_t i = i0 + threadIdx.x;
    if (i < i1) {
        uint8_t rr = 0;
        for (int64_t q = img_off[i]; q < img_off[i + 1]; ++q) {
            int32_t v = label_id[q];
            if (v >= 0 && v < n_vocab && __ldg(lut_nrep + v) > 0) { rr = 1; break; }
        }
        row_rep[i] = rr; c_rrow += rr;
    }
    // block reduction -> six global atomics per block
    __shared__ unsigned long long red[6][LB_THREADS / 32];
    unsigned long long vals[6] = {c_obj, c_miss, c_lab, c_rlab, c_robj, c_rrow};
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        unsigned long long v = vals[k];
        for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(FULL, v, off);
        if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = v;
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        unsigned long long s = 0;
        for (int w = 0; w < LB_THREADS / 32; ++w) s += red[threadIdx.x][w];
        if (s) atomicAdd(&counters[threadIdx.x], s);
    }
}

// per-vocabulary occurrence counts (the unmatched-label table of processor.py:591-593 is derived
// from them on the host); shared-memory privatised for vocabularies that fit, global atomics otherwise
constexpr int HIST_SMEM = 4096;
__global__ void __launch_bounds__(LB_THREADS)
label_hist_kernel(const int32_t* __restrict__ label_id, int64_t n_box, int32_t n_vocab, unsigned long long* hist) {
    __shared__ unsigned sh[HIST_SMEM];
    const bool priv = n_vocab <= HIST_SMEM;
    if (priv) for (int v = threadIdx.x; v < n_vocab; v += LB_THREADS) sh[v] = 0;
    __syncthreads();
    auto add = [&](int32_t v) {
        if (v < 0 || v >= n_vocab) return;
        if (priv) atomicAdd(&sh[v], 1u); else atomicAdd(&hist[v], 1ULL);
    };
    const int64_t per = (((n_box + gridDim.x - 1) / gridDim.x) + 3) & ~(int64_t)3;     // multiple of 4: slices stay aligned
    const int64_t a = min(blockIdx.x * per, n_box), b = min(a + per, n_box);
    const bool vec = (reinterpret_cast<uintptr_t>(label_id) & 15) == 0;
    const int64_t bv = vec ? a + ((b - a) & ~(int64_t)3) : a;
    for (int64_t q = a + 4 * (int64_t)threadIdx.x; q < bv; q += 4 * LB_THREADS) {
        const int4 v = *reinterpret_cast<const int4*>(label_id + q);
        add(v.x); add(v.y); add(v.z); add(v.w);
    }
    for (int64_t q = bv + threadIdx.x; q < b; q += LB_THREADS) add(label_id[q]);
    __syncthreads();
    if (priv) for (int v = threadIdx.x; v < n_vocab; v += LB_THREADS) if (sh[v]) atomicAdd(&hist[v], (unsigned long long)sh[v]);
}

// ------------------------------------------------------------------------------- K6
__device__ __forceinline__ int category_of(const int32_t* __restrict__ label_id, const int32_t* __restrict__ cat_of_label,
                                           int32_t n_vocab, int64_t q) {
    int32_t v = label_id[q];
    if (v < 0 || v >= n_vocab) return -1;
    return __ldg(cat_of_label + v);
}

// counts[chunk][cat] for chunk = IMG_PER_WARP consecutive images, one warp per chunk
__global__ void __launch_bounds__(LB_THREADS)
split_count_kernel(const int64_t* __restrict__ img_off, int64_t n_img, const int32_t* __restrict__ label_id,
                   const int32_t* __restrict__ cat_of_label, int32_t n_vocab, int32_t n_cat,
                   unsigned long long* __restrict__ counts, int64_t n_chunks) {
    __shared__ unsigned cnt[LB_THREADS / 32][MAX_CAT];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t chunk = blockIdx.x * (int64_t)(LB_THREADS / 32) + w;
    for (int c = lane; c < n_cat; c += 32) cnt[w][c] = 0;
    __syncwarp();
    if (chunk < n_chunks) {
        const int64_t i0 = chunk * IMG_PER_WARP, i1 = min(i0 + IMG_PER_WARP, n_img);
        const int64_t q0 = img_off[i0], q1 = img_off[i1];
        for (int64_t q = q0 + lane; q < q1; q += 32) {
            int c = category_of(label_id, cat_of_label, n_vocab, q);
            if (c >= 0 && c < n_cat) atomicAdd(&cnt[w][c], 1u);
        }
        __syncwarp();
        for (int c = lane; c < n_cat; c += 32) counts[chunk * n_cat + c] = cnt[w][c];
    }
}

// per category: exclusive scan of counts over chunks (in place) + total; one block per category
__global__ void __launch_bounds__(1024)
split_scan_kernel(unsigned long long* counts, int64_t n_chunks, int32_t n_cat, unsigned long long* totals) {
    __shared__ unsigned long long wsum[32];
    __shared__ unsigned long long carry;
    const int c = blockIdx.x;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < n_chunks; base += 1024) {
        const int64_t k = base + threadIdx.x;
        unsigned long long v = k < n_chunks ? counts[k * n_cat + c] : 0ULL;
        unsigned long long x = v;
        for (int off = 1; off < 32; off <<= 1) { unsigned long long y = __shfl_up_sync(FULL, x, off); if (lane >= off) x += y; }
        if (lane == 31) wsum[w] = x;
        __syncthreads();
        if (w == 0) {
            unsigned long long s = wsum[lane], t = s;
            for (int off = 1; off < 32; off <<= 1) { unsigned long long y = __shfl_up_sync(FULL, t, off); if (lane >= off) t += y; }
            wsum[lane] = t - s;                       // exclusive warp offsets
        }
        __syncthreads();
        const unsigned long long excl = carry + wsum[w] + (x - v);
        if (k < n_chunks) counts[k * n_cat + c] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) totals[c] = carry;
}

__global__ void split_catoff_kernel(const unsigned long long* __restrict__ totals, int32_t n_cat, int64_t* __restrict__ cat_off) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        int64_t acc = 0;
        cat_off[0] = 0;
        for (int c = 0; c < n_cat; ++c) { acc += (int64_t)totals[c]; cat_off[c + 1] = acc; }
    }
}

// stable scatter: a warp walks its chunk 32 objects at a time; objects of one category keep
// their (image, object) order because ranks follow lane order and chunks follow image order
__global__ void __launch_bounds__(LB_THREADS)
split_fill_kernel(const int64_t* __restrict__ img_off, int64_t n_img, const int32_t* __restrict__ label_id,
                  const int32_t* __restrict__ cat_of_label, int32_t n_vocab, int32_t n_cat,
                  const int64_t* __restrict__ cat_off, const unsigned long long* __restrict__ chunk_excl, int64_t n_chunks,
                  int64_t* __restrict__ exp_img, int64_t* __restrict__ exp_box, int32_t* __restrict__ exp_cat) {
    __shared__ unsigned long long cur[LB_THREADS / 32][MAX_CAT];
    __shared__ int64_t soff[LB_THREADS / 32][IMG_PER_WARP + 1];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t chunk = blockIdx.x * (int64_t)(LB_THREADS / 32) + w;
    if (chunk >= n_chunks) return;
    const int64_t i0 = chunk * IMG_PER_WARP, i1 = min(i0 + IMG_PER_WARP, n_img);
    const int ni = (int)(i1 - i0);
    for (int c = lane; c < n_cat; c += 32) cur[w][c] = (unsigned long long)cat_off[c] + chunk_excl[chunk * n_cat + c];
    for (int k = lane; k <= ni; k += 32) soff[w][k] = img_off[i0 + k];
    __syncwarp();
    const int64_t q0 = soff[w][0], q1 = soff[w][ni];
    for (int64_t base = q0; base < q1; base += 32) {
        const int64_t q = base + lane;
        int c = -1;
        if (q < q1) { c = category_of(label_id, cat_of_label, n_vocab, q); if (c >= n_cat) c = -1; }
        const unsigned peers = __match_any_sync(FULL, c);
        if (c >= 0) {
            const int rank = __popc(peers & ((1u << lane) - 1));
            const unsigned long long dst = cur[w][c] + rank;
            // image of object q: largest k with soff[k] <= q
            int lo = 0, hi = ni;
            while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (soff[w][mid] <= q) lo = mid; else hi = mid; }
            exp_img[dst] = i0 + lo; exp_box[dst] = q; exp_cat[dst] = c;
        }
        __syncwarp();
        if (c >= 0 && lane == __ffs(peers) - 1) cur[w][c] += __popc(peers);
        __syncwarp();
    }
}

// The same scatter with the rows of each category staged in shared memory and written out 32 at a
// time: three fully coalesced stores per 32 rows instead of a few 8-byte pieces per step (the plain
// kernel spends 70 % of its time in those partial-sector writes).  Up to ST_CAT categories.
constexpr int ST_CAT = 16, ST_WARPS = 4, ST_RING = 64;
__global__ void __launch_bounds__(32 * ST_WARPS)
split_fill_staged_kernel(const int64_t* __restrict__ img_off, int64_t n_img, const int32_t* __restrict__ label_id,
                         const int32_t* __restrict__ cat_of_label, int32_t n_vocab, int32_t n_cat,
                         const int64_t* __restrict__ cat_off, const unsigned long long* __restrict__ chunk_excl, int64_t n_chunks,
                         int64_t* __restrict__ exp_img, int64_t* __restrict__ exp_box, int32_t* __restrict__ exp_cat) {
    __shared__ uint2 ring[ST_WARPS][ST_CAT][ST_RING];          // (image - i0, object - q0) of rows waiting to be written
    __shared__ unsigned long long cur[ST_WARPS][ST_CAT];       // next global row of each category
    __shared__ int fill[ST_WARPS][ST_CAT];
    __shared__ int64_t soff[ST_WARPS][IMG_PER_WARP + 1];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t chunk = blockIdx.x * (int64_t)ST_WARPS + w;
    if (chunk >= n_chunks) return;
    const int64_t i0 = chunk * IMG_PER_WARP, i1 = min(i0 + IMG_PER_WARP, n_img);
    const int ni = (int)(i1 - i0);
    if (lane < n_cat) { cur[w][lane] = (unsigned long long)cat_off[lane] + chunk_excl[chunk * n_cat + lane]; fill[w][lane] = 0; }
    for (int k = lane; k <= ni; k += 32) soff[w][k] = img_off[i0 + k];
    __syncwarp();
    const int64_t q0 = soff[w][0], q1 = soff[w][ni];
    auto flush = [&](int c, int count) {                       // writes ring[c][0 .. count) (count <= 32), keeps the rest
        const unsigned long long dst = cur[w][c];
        const int have = fill[w][c];
        if (lane < count) {
            const uint2 e = ring[w][c][lane];
            exp_img[dst + lane] = i0 + e.x; exp_box[dst + lane] = q0 + e.y; exp_cat[dst + lane] = c;
        }
        uint2 keep = make_uint2(0, 0);
        if (count + lane < have) keep = ring[w][c][count + lane];
        __syncwarp();
        if (count + lane < have) ring[w][c][lane] = keep;
        if (lane == 0) { cur[w][c] = dst + count; fill[w][c] = have - count; }
        __syncwarp();
    };
    for (int64_t base = q0; base < q1; base += 32) {
        const int64_t q = base + lane;
        int c = -1;
        if (q < q1) { c = category_of(label_id, cat_of_label, n_vocab, q); if (c >= n_cat) c = -1; }
        const unsigned peers = __match_any_sync(FULL, c);
        bool leader = false;
        if (c >= 0) {
            const int rank = __popc(peers & ((1u << lane) - 1));
            int lo = 0, hi = ni;                               // image of object q: largest k with soff[k] <= q
            while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (soff[w][mid] <= q) lo = mid; else hi = mid; }
            ring[w][c][fill[w][c] + rank] = make_uint2((unsigned)lo, (unsigned)(q - q0));
            leader = lane == __ffs(peers) - 1;
        }
        __syncwarp();
        int now = 0;
        if (leader) { now = fill[w][c] + __popc(peers); fill[w][c] = now; }
        unsigned full = __reduce_or_sync(FULL, (leader && now >= 32) ? (1u << c) : 0u);
        __syncwarp();
        for (; full; full &= full - 1) flush(__ffs(full) - 1, 32);
    }
    for (int c = 0; c < n_cat; ++c) {
        const int have = fill[w][c];                           // < 32 here
        if (have > 0) flush(c, have);
    }
}

__global__ void __launch_bounds__(LB_THREADS)
split_assign_kernel(const int64_t* __restrict__ cat_off, int32_t n_cat, const int64_t* __restrict__ perm, int64_t n_exp,
                    const int64_t* __restrict__ n_train, const int64_t* __restrict__ n_val,
                    uint8_t* __restrict__ split, int64_t* __restrict__ pos) {
    const int64_t r = blockIdx.x * (int64_t)LB_THREADS + threadIdx.x;   // position in the shuffled order
    if (r >= n_exp) return;
    int lo = 0, hi = n_cat;
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (cat_off[mid] <= r) lo = mid; else hi = mid; }
    const int64_t base = cat_off[lo], rp = r - base;
    const int64_t j = perm[r];                                           // original row taken to position rp
    const int64_t ntr = n_train[lo], nva = n_val[lo];
    split[base + j] = rp < ntr ? 0 : (rp < ntr + nva ? 1 : 2);
    pos[base + j] = rp;
}

// Sharded form: the permutation covers the GLOBAL category-grouped table, this rank owns, per category c, the rows
// own_lo[c] .. own_lo[c] + own_cnt[c] of the category (sharding.split_category_bases) and keeps them at local_off[c] .. in
// its own arrays.  Every rank sweeps the whole permutation (coalesced 16-byte loads) and keeps what lands in its ranges.
__global__ void __launch_bounds__(LB_THREADS)
split_assign_range_kernel(const int64_t* __restrict__ cat_off, int32_t n_cat, const int64_t* __restrict__ perm, int64_t n_exp,
                          const int64_t* __restrict__ n_train, const int64_t* __restrict__ n_val,
                          const int64_t* __restrict__ own_lo, const int64_t* __restrict__ own_cnt, const int64_t* __restrict__ local_off,
                          uint8_t* __restrict__ split, int64_t* __restrict__ pos) {
    const int64_t r0 = (blockIdx.x * (int64_t)LB_THREADS + threadIdx.x) * 2;   // two shuffled positions per thread
    if (r0 >= n_exp) return;
    long long j2[2];
    if (r0 + 1 < n_exp) { const longlong2 v = __ldg(reinterpret_cast<const longlong2*>(perm + r0)); j2[0] = v.x; j2[1] = v.y; }
    else { j2[0] = perm[r0]; j2[1] = -1; }
    int lo = 0, hi = n_cat;
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (cat_off[mid] <= r0) lo = mid; else hi = mid; }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int64_t r = r0 + u;
        if (r >= n_exp) break;
        while (lo + 1 < n_cat && cat_off[lo + 1] <= r) ++lo;             // the second position may start the next category
        const int64_t rp = r - cat_off[lo];
        const int64_t k = j2[u] - own_lo[lo];
        if (k >= 0 && k < own_cnt[lo]) {
            const int64_t ntr = n_train[lo], nva = n_val[lo];
            const int64_t at = local_off[lo] + k;
            split[at] = rp < ntr ? 0 : (rp < ntr + nva ? 1 : 2);
            pos[at] = rp;
        }
    }
}

// ------------------------------------------------------------------------------- YOLO normalisation
// processor.py:1045-1052.  A block owns IMG_PER_CTA consecutive images = one contiguous object range:
// the image offsets and sizes go to shared memory, then one thread per object streams its corner
// points in and its (cx, cy, w, h) out with 128-bit accesses and finds its image by bisection.
__global__ void __launch_bounds__(LB_THREADS)
yolo_kernel(const int64_t* __restrict__ img_off, const double* __restrict__ pts, const uint8_t* __restrict__ valid,
            const double* __restrict__ img_wh, int64_t n_img, double* __restrict__ out, uint8_t* __restrict__ ok) {
    __shared__ long long soff[IMG_PER_CTA + 1];
    __shared__ double2 swh[IMG_PER_CTA];
    const int64_t i0 = blockIdx.x * (int64_t)IMG_PER_CTA;
    const int n_here = (int)min((int64_t)IMG_PER_CTA, n_img - i0);
    for (int k = threadIdx.x; k <= n_here; k += LB_THREADS) soff[k] = img_off[i0 + k];
    for (int k = threadIdx.x; k < n_here; k += LB_THREADS) swh[k] = reinterpret_cast<const double2*>(img_wh)[i0 + k];
    __syncthreads();
    const int64_t q0 = soff[0], q1 = soff[n_here];
    const double2* pts2 = reinterpret_cast<const double2*>(pts);
    double2* out2 = reinterpret_cast<double2*>(out);
    for (int64_t q = q0 + threadIdx.x; q < q1; q += LB_THREADS) {
        int lo = 0, hi = n_here;                   // last image whose first object is <= q (skips empty images)
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (soff[mid] <= q) lo = mid; else hi = mid; }
        const double W = swh[lo].x, H = swh[lo].y;
        const double2 p1 = ldg_stream_f64x2(pts2 + 2 * q), p2 = ldg_stream_f64x2(pts2 + 2 * q + 1);
        const double x1 = pymin(p1.x, p2.x), x2 = pymax(p1.x, p2.x), y1 = pymin(p1.y, p2.y), y2 = pymax(p1.y, p2.y);
        const double bw = pymax(__dsub_rn(x2, x1), 0.0), bh = pymax(__dsub_rn(y2, y1), 0.0);
        const bool good = (valid == nullptr || valid[q]) && !(bw <= 0.0) && !(bh <= 0.0) && W != 0.0 && H != 0.0;
        ok[q] = good ? 1 : 0;
        stg_stream_f64x2(out2 + 2 * q, make_double2(__ddiv_rn(__ddiv_rn(__dadd_rn(x1, x2), 2.0), W),
                                                    __ddiv_rn(__ddiv_rn(__dadd_rn(y1, y2), 2.0), H)));
        stg_stream_f64x2(out2 + 2 * q + 1, make_double2(__ddiv_rn(bw, W), __ddiv_rn(bh, H)));
    }
}

// ------------------------------------------------------------------------------- per-image label presence
// summarize_yolo_label_counts (processor.py:1113-1131): a label counts once per image that holds it ("图片数量") and once
// per box ("标注框数量").  A block owns IMG_PER_CTA consecutive images; one thread per object finds its image by bisection
// and adds to the image histogram iff no earlier object of the same image carries the same label (images hold tens of
// boxes, so the backward scan is short and L1-resident).  Ids outside [0, n_vocab) are ignored.
__global__ void __launch_bounds__(LB_THREADS)
label_presence_kernel(const int64_t* __restrict__ img_off, const int32_t* __restrict__ label_id, int64_t n_img, int32_t n_vocab,
                      unsigned long long* __restrict__ img_hist, unsigned long long* __restrict__ box_hist) {
    __shared__ long long soff[IMG_PER_CTA + 1];
    const int64_t i0 = blockIdx.x * (int64_t)IMG_PER_CTA;
    const int n_here = (int)min((int64_t)IMG_PER_CTA, n_img - i0);
    for (int k = threadIdx.x; k <= n_here; k += LB_THREADS) soff[k] = img_off[i0 + k];
    __syncthreads();
    const int64_t q0 = soff[0], q1 = soff[n_here];
    for (int64_t q = q0 + threadIdx.x; q < q1; q += LB_THREADS) {
        const int32_t v = label_id[q];
        if (v < 0 || v >= n_vocab) continue;
        int lo = 0, hi = n_here;                   // last image whose first object is <= q (skips empty images)
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (soff[mid] <= q) lo = mid; else hi = mid; }
        bool first = true;
        for (int64_t e = soff[lo]; e < q && first; ++e) first = label_id[e] != v;
        atomicAdd(&box_hist[v], 1ULL);
        if (first) atomicAdd(&img_hist[v], 1ULL);
    }
}

static inline int64_t n_chunks_of(int64_t n_img) { return (n_img + IMG_PER_WARP - 1) / IMG_PER_WARP; }

}  // namespace dyd

using namespace dyd;

extern "C" int dyd_label_lut(const int64_t* d_img_off, const int32_t* d_label_id, int64_t n_img, int64_t n_box,
                             const int32_t* d_lut_new, const int32_t* d_lut_ntok, const int32_t* d_lut_nrep,
                             int32_t n_vocab, int32_t* d_new_id, uint8_t* d_row_replaced, uint64_t* d_counters, void* stream) {
    DYD_REQUIRE(n_img >= 0 && n_box >= 0 && n_vocab >= 0, DYD_E_ARG, "negative count");
    DYD_REQUIRE(d_counters, DYD_E_ARG, "null pointer");
    cudaStream_t s = as_stream(stream);
    DYD_CUDA(cudaMemsetAsync(d_counters, 0, 6 * sizeof(uint64_t), s));
    if (n_img == 0) return 0;
    DYD_REQUIRE(d_img_off && d_row_replaced && (n_box == 0 || (d_label_id && d_new_id)) &&
                    (n_vocab == 0 || (d_lut_new && d_lut_ntok && d_lut_nrep)), DYD_E_ARG, "null pointer");
    const int64_t grid = (n_img + IMG_PER_CTA - 1) / IMG_PER_CTA;
    label_lut_kernel<<<(unsigned)grid, LB_THREADS, 0, s>>>(d_img_off, d_label_id, n_img, d_lut_new, d_lut_ntok, d_lut_nrep,
                                                           n_vocab, d_new_id, d_row_replaced,
                                                           reinterpret_cast<unsigned long long*>(d_counters));
    return launch_check("label_lut_kernel");
}

extern "C" int dyd_label_hist(const int32_t* d_label_id, int64_t n_box, int32_t n_vocab, uint64_t* d_hist, void* stream) {
    DYD_REQUIRE(n_box >= 0 && n_vocab >= 0, DYD_E_ARG, "negative count");
    if (n_vocab == 0) return 0;
    DYD_REQUIRE(d_hist && (n_box == 0 || d_label_id), DYD_E_ARG, "null pointer");
    cudaStream_t s = as_stream(stream);
    DYD_CUDA(cudaMemsetAsync(d_hist, 0, sizeof(uint64_t) * n_vocab, s));
    if (n_box == 0) return 0;
    // each block covers at most ~1M objects so the 32-bit private counters cannot overflow
    int64_t grid = (n_box + (1 << 20) - 1) >> 20;
    if (grid < NUM_SMS * 2) grid = NUM_SMS * 2;
    label_hist_kernel<<<(unsigned)grid, LB_THREADS, 0, s>>>(d_label_id, n_box, n_vocab, reinterpret_cast<unsigned long long*>(d_hist));
    return launch_check("label_hist_kernel");
}

extern "C" int dyd_label_presence(const int64_t* d_img_off, const int32_t* d_label_id, int64_t n_img, int32_t n_vocab,
                                  uint64_t* d_img_hist, uint64_t* d_box_hist, void* stream) {
    DYD_REQUIRE(n_img >= 0 && n_vocab >= 0, DYD_E_ARG, "negative count");
    if (n_vocab == 0) return 0;
    DYD_REQUIRE(d_img_hist && d_box_hist, DYD_E_ARG, "null pointer");
    cudaStream_t s = as_stream(stream);
    DYD_CUDA(cudaMemsetAsync(d_img_hist, 0, sizeof(uint64_t) * n_vocab, s));
    DYD_CUDA(cudaMemsetAsync(d_box_hist, 0, sizeof(uint64_t) * n_vocab, s));
    if (n_img == 0) return 0;
    DYD_REQUIRE(d_img_off && d_label_id, DYD_E_ARG, "null pointer");
    const int64_t grid = (n_img + IMG_PER_CTA - 1) / IMG_PER_CTA;
    label_presence_kernel<<<(unsigned)grid, LB_THREADS, 0, s>>>(d_img_off, d_label_id, n_img, n_vocab,
                                                               reinterpret_cast<unsigned long long*>(d_img_hist),
                                                               reinterpret_cast<unsigned long long*>(d_box_hist));
    return launch_check("label_presence_kernel");
}

extern "C" size_t dyd_split_workspace_bytes(int64_t n_img, int32_t n_cat) {
    if (n_img < 0) n_img = 0;
    if (n_cat < 0) n_cat = 0;
    return ((size_t)n_chunks_of(n_img) * (size_t)n_cat + (size_t)n_cat + 2) * sizeof(unsigned long long);
}

extern "C" int dyd_split_count(const int64_t* d_img_off, int64_t n_img, const int32_t* d_label_id,
                               const int32_t* d_cat_of_label, int32_t n_vocab, int32_t n_cat, int64_t* d_cat_off,
                               void* ws, size_t ws_bytes, void* stream) {
    DYD_REQUIRE(n_img >= 0 && n_vocab >= 0 && n_cat >= 0 && n_cat <= MAX_CAT, DYD_E_ARG, "bad count (n_cat <= 256)");
    DYD_REQUIRE(d_cat_off && ws, DYD_E_ARG, "null pointer");
    DYD_REQUIRE(ws_bytes >= dyd_split_workspace_bytes(n_img, n_cat), DYD_E_WORKSPACE, "workspace too small");
    cudaStream_t s = as_stream(stream);
    const int64_t nch = n_chunks_of(n_img);
    unsigned long long* counts = reinterpret_cast<unsigned long long*>(ws);
    unsigned long long* totals = counts + (size_t)nch * n_cat;
    if (n_img == 0 || n_cat == 0) { DYD_CUDA(cudaMemsetAsync(d_cat_off, 0, (n_cat + 1) * sizeof(int64_t), s)); return 0; }
    DYD_REQUIRE(d_img_off && d_label_id && d_cat_of_label, DYD_E_ARG, "null pointer");
    const int64_t grid = (nch + LB_THREADS / 32 - 1) / (LB_THREADS / 32);
    split_count_kernel<<<(unsigned)grid, LB_THREADS, 0, s>>>(d_img_off, n_img, d_label_id, d_cat_of_label, n_vocab, n_cat, counts, nch);
    if (int rc = launch_check("split_count_kernel")) return rc;
    split_scan_kernel<<<n_cat, 1024, 0, s>>>(counts, nch, n_cat, totals);
    if (int rc = launch_check("split_scan_kernel")) return rc;
    split_catoff_kernel<<<1, 32, 0, s>>>(totals, n_cat, d_cat_off);
    return launch_check("split_catoff_kernel");
}

extern "C" int dyd_split_fill(const int64_t* d_img_off, int64_t n_img, const int32_t* d_label_id,
                              const int32_t* d_cat_of_label, int32_t n_vocab, int32_t n_cat, const int64_t* d_cat_off,
                              int64_t* d_exp_img, int64_t* d_exp_box, int32_t* d_exp_cat,
                              void* ws, size_t ws_bytes, void* stream) {
    DYD_REQUIRE(n_img >= 0 && n_vocab >= 0 && n_cat >= 0 && n_cat <= MAX_CAT, DYD_E_ARG, "bad count (n_cat <= 256)");
    if (n_img == 0 || n_cat == 0) return 0;
    DYD_REQUIRE(d_img_off && d_label_id && d_cat_of_label && d_cat_off && d_exp_img && d_exp_box && d_exp_cat && ws, DYD_E_ARG, "null pointer");
    DYD_REQUIRE(ws_bytes >= dyd_split_workspace_bytes(n_img, n_cat), DYD_E_WORKSPACE, "workspace too small");
    const int64_t nch = n_chunks_of(n_img);
    if (n_cat <= ST_CAT) {
        const int64_t grid = (nch + ST_WARPS - 1) / ST_WARPS;
        split_fill_staged_kernel<<<(unsigned)grid, 32 * ST_WARPS, 0, as_stream(stream)>>>(
            d_img_off, n_img, d_label_id, d_cat_of_label, n_vocab, n_cat, d_cat_off,
            reinterpret_cast<const unsigned long long*>(ws), nch, d_exp_img, d_exp_box, d_exp_cat);
        return launch_check("split_fill_staged_kernel");
    }
    const int64_t grid = (nch + LB_THREADS / 32 - 1) / (LB_THREADS / 32);
    split_fill_kernel<<<(unsigned)grid, LB_THREADS, 0, as_stream(stream)>>>(
        d_img_off, n_img, d_label_id, d_cat_of_label, n_vocab, n_cat, d_cat_off,
        reinterpret_cast<const unsigned long long*>(ws), nch, d_exp_img, d_exp_box, d_exp_cat);
    return launch_check("split_fill_kernel");
}

extern "C" int dyd_split_assign(const int64_t* d_cat_off, int32_t n_cat, const int64_t* d_perm, int64_t n_exp,
                                const int64_t* d_n_train, const int64_t* d_n_val, uint8_t* d_split, int64_t* d_pos, void* stream) {
    DYD_REQUIRE(n_exp >= 0 && n_cat >= 0, DYD_E_ARG, "negative count");
    if (n_exp == 0) return 0;
    DYD_REQUIRE(d_cat_off && d_perm && d_n_train && d_n_val && d_split && d_pos && n_cat > 0, DYD_E_ARG, "null pointer");
    const int64_t grid = (n_exp + LB_THREADS - 1) / LB_THREADS;
    split_assign_kernel<<<(unsigned)grid, LB_THREADS, 0, as_stream(stream)>>>(d_cat_off, n_cat, d_perm, n_exp, d_n_train, d_n_val, d_split, d_pos);
    return launch_check("split_assign_kernel");
}

extern "C" int dyd_split_assign_range(const int64_t* d_cat_off, int32_t n_cat, const int64_t* d_perm, int64_t n_exp,
                                      const int64_t* d_n_train, const int64_t* d_n_val, const int64_t* d_own_lo,
                                      const int64_t* d_own_cnt, const int64_t* d_local_off, uint8_t* d_split, int64_t* d_pos, void* stream) {
    DYD_REQUIRE(n_exp >= 0 && n_cat >= 0, DYD_E_ARG, "negative count");
    if (n_exp == 0) return 0;
    DYD_REQUIRE(d_cat_off && d_perm && d_n_train && d_n_val && d_own_lo && d_own_cnt && d_local_off && d_split && d_pos && n_cat > 0, DYD_E_ARG,
                "null pointer");
    DYD_REQUIRE(((uintptr_t)d_perm & 15) == 0, DYD_E_ALIGN, "perm must be 16-byte aligned");
    const int64_t grid = ((n_exp + 1) / 2 + LB_THREADS - 1) / LB_THREADS;
    split_assign_range_kernel<<<(unsigned)grid, LB_THREADS, 0, as_stream(stream)>>>(d_cat_off, n_cat, d_perm, n_exp, d_n_train, d_n_val, d_own_lo,
                                                                                   d_own_cnt, d_local_off, d_split, d_pos);
    return launch_check("split_assign_range_kernel");
}

extern "C" int dyd_yolo_normalise(const int64_t* d_img_off, const double* d_pts, const uint8_t* d_valid,
                                  const double* d_img_wh, int64_t n_img, int64_t n_box,
                                  double* d_cxcywh, uint8_t* d_ok, void* stream) {
    DYD_REQUIRE(n_img >= 0 && n_box >= 0, DYD_E_ARG, "negative count");
    if (n_img == 0 || n_box == 0) return 0;
    DYD_REQUIRE(d_img_off && d_pts && d_img_wh && d_cxcywh && d_ok, DYD_E_ARG, "null pointer");
    DYD_REQUIRE(((uintptr_t)d_pts & 15) == 0 && ((uintptr_t)d_cxcywh & 15) == 0 && ((uintptr_t)d_img_wh & 15) == 0, DYD_E_ALIGN,
                "pts / img_wh / cxcywh must be 16-byte aligned");
    const int64_t grid = (n_img + IMG_PER_CTA - 1) / IMG_PER_CTA;
    yolo_kernel<<<(unsigned)grid, LB_THREADS, 0, as_stream(stream)>>>(d_img_off, d_pts, d_valid, d_img_wh, n_img, d_cxcywh, d_ok);
    return launch_check("yolo_kernel");
}

// K2 on one tile: up to 32 boxes of up to K2_TILE_IMAGES consecutive images sit in shared memory
// (two arrays so that consecutive boxes are consecutive 16-byte bank groups); one warp decides
// "any pair of an image reaches the threshold" for all of them at once
// (processor.py:368-376 meet_conditions, :341-366 extract_boxes' prefix rule, :328-339 calculate_iou).
// Shared by the fused kernel (bbox_tma.cu) and the stand-alone K2 kernel (bbox_iou.cu).
#pragma once
#include "bbox_core.cuh"

namespace dyd {

constexpr int K2_QCAP = 64;                    // survivor queue entries per warp

struct K2Tile {                                // per-warp shared memory
    double2 box_lo[32], box_hi[32];            // (x1, y1) / (x2, y2) after extract_boxes' min/max
    int lq[8];                                 // per image: first object (64 past the last image)
    unsigned short queue[K2_QCAP];             // pairs that passed the overlap pre-test: a | b << 6 | image << 12
};

__device__ __forceinline__ Box k2_load_box(const K2Tile& t, int q) {
    const double2 lo = t.box_lo[q], hi = t.box_hi[q];
    return Box{lo.x, lo.y, hi.x, hi.y};
}
// Exact overlap pre-test of calculate_iou (processor.py:329-335): the pair can only reach the
// threshold if both max(0, .) terms are positive.  Same selects and subtractions as iou_hits.
__device__ __forceinline__ bool boxes_overlap(const Box& a, const Box& b) {
    const double xi1 = pymax(a.x1, b.x1), yi1 = pymax(a.y1, b.y1);
    const double xi2 = pymin(a.x2, b.x2), yi2 = pymin(a.y2, b.y2);
    return __dsub_rn(xi2, xi1) > 0.0 && __dsub_rn(yi2, yi1) > 0.0;
}
static __device__ __noinline__ bool iou_hits_cold(const Box& a, const Box& b, double thr, bool zero_hits) {
    return iou_hits(a, b, thr, zero_hits);
}

// All 32 lanes call.  Boxes 0..np-1 (np <= 32) are in t.box_*; bit p of `inv` marks a null bbox;
// lane j < ni describes image j of the tile: my_a = its first object (tile-local), my_n = its object
// count (lanes >= ni pass anything).  TM = most images a tile can hold (<= 8).  Returns the hit mask
// (bit j: image j is high-IoU) in every lane and, in lane j < ni, image j's box count in my_ne.
// Leaves t.lq / t.queue in use until the caller's next __syncwarp().
template <int TM>
__device__ __forceinline__ unsigned k2_tile_any(K2Tile& t, unsigned inv, bool exact_pre, int my_a, int my_n, int ni, int np,
                                                int64_t min_boxes, double thr, bool zero_hits, int lane, int& my_ne) {
    static_assert(TM <= 8, "lq holds 8 images");
    // lane j < ni owns image j: its box count is the prefix before the first null bbox
    int my_act = 0;
    my_ne = 0;
    if (lane < ni) {
        const unsigned m = my_n >= 32 ? (inv >> my_a) : ((inv >> my_a) & ((1u << my_n) - 1u));
        my_ne = m ? __ffs(m) - 1 : my_n;
        if (my_ne >= min_boxes && my_ne >= 2) my_act = my_ne;
    } else {
        my_a = 64;
    }
    if (lane < 8) t.lq[lane] = my_a;
    __syncwarp();
    // Lane b owns box b and meets the boxes of its own image at circular distance d = 1 .. n/2 (every
    // unordered pair exactly once).  Pass 1 is an overlap pre-test: without NaN coordinates the two
    // cross comparisons per axis are a superset of the exact test, which the survivors get anyway; a
    // tile holding a NaN box uses the reference's selects in the reference's argument order (lower
    // index first).  Pass 2 runs the full arithmetic on the survivors (a few per cent).
    int j = 0;                                 // image of box `lane`: the last one starting at or before it
#pragma unroll
    for (int k = 1; k < TM; ++k) j += lane >= t.lq[k] ? 1 : 0;
    const int first = __shfl_sync(FULL, my_a, j), ne = __shfl_sync(FULL, my_act, j);
    const int a = lane - first;
    const int half = (lane < np && a < ne) ? ne >> 1 : 0;
    const int maxhalf = __reduce_max_sync(FULL, half);
    const Box mine = k2_load_box(t, lane);     // lanes that own no box read stale data and never use it
    unsigned sv = 0;                           // bit d: the pair at distance d passed the pre-test
    for (int d = 1; d <= maxhalf; ++d) {
        const bool on = d <= half && (2 * d != ne || a < d);          // even n: distance n/2 pairs appear twice
        int pb = a + d;
        pb = pb >= ne ? pb - ne : pb;
        const int ib = on ? first + pb : lane;
        const Box o = k2_load_box(t, ib);
        bool ov;
        if (exact_pre) ov = ib > lane ? boxes_overlap(mine, o) : boxes_overlap(o, mine);
        else ov = mine.x2 > o.x1 && o.x2 > mine.x1 && mine.y2 > o.y1 && o.y2 > mine.y1;
        if (on && (zero_hits || ov)) sv |= 1u << d;
    }
    unsigned hits = 0;
    if (__any_sync(FULL, sv != 0)) {
        // Survivors sit unevenly in the lanes: spread them over a small queue so that the full
        // arithmetic runs on dense warps.
        const int cnt = __popc(sv);
        int incl = cnt;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) { const int y = __shfl_up_sync(FULL, incl, off); if (lane >= off) incl += y; }
        const int total = __shfl_sync(FULL, incl, 31);
        if (total <= K2_QCAP) {
            int pos = incl - cnt;
            for (unsigned m = sv; m; m &= m - 1) {
                const int d = __ffs(m) - 1;
                int pb = a + d;
                pb = pb >= ne ? pb - ne : pb;
                const int ib = first + pb;
                t.queue[pos++] = (unsigned short)(min(lane, ib) | (max(lane, ib) << 6) | (j << 12));
            }
            __syncwarp();
            for (int base = 0; base < total; base += 32) {
                bool hit = false; int jj = 0;
                if (base + lane < total) {
                    const unsigned e = t.queue[base + lane];
                    jj = e >> 12;
                    hit = iou_hits(k2_load_box(t, e & 63), k2_load_box(t, (e >> 6) & 63), thr, zero_hits);
                }
                hits |= __reduce_or_sync(FULL, hit ? (1u << jj) : 0u);
            }
        } else {                               // crowded tile: every lane works through its own pairs
            while (__any_sync(FULL, sv != 0)) {
                bool hit = false;
                if (sv) {
                    const int d = __ffs(sv) - 1;
                    sv &= sv - 1;
                    int pb = a + d;
                    pb = pb >= ne ? pb - ne : pb;
                    const int ib = first + pb;
                    const Box o = k2_load_box(t, ib);
                    // without NaN the arithmetic is symmetric in its arguments bit for bit
                    hit = (exact_pre && ib < lane) ? iou_hits_cold(o, mine, thr, zero_hits) : iou_hits(mine, o, thr, zero_hits);
                }
                hits |= __reduce_or_sync(FULL, hit ? (1u << j) : 0u);
                if ((hits >> j) & 1u) sv = 0;  // any() is settled for this image
            }
        }
    }
    return hits;
}

}  // namespace dyd

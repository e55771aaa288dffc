// K1 (polygon -> corner points), K2 (box count + any-pair IoU) and the direct-load fused path.
//
// Direct-load kernels: warps read the vertex stream with 128-bit non-allocating loads, a
// G-lane group per polygon (32/G polygons per warp step).  The fused kernel keeps each image's
// boxes in shared memory for the pair test, so boxes are written to HBM once and never
// re-read.  The stand-alone K2 kernel packs consecutive images into 32-box tiles (one lane per
// box, k2_tile.cuh).  Crowded images (more than 32 objects for K2, more than WARP_BOX_CAP for the
// direct fused kernel) go to a worklist and are finished by a block-per-image kernel
// (shared-memory SoA tile, circular half-range pairing).
// The TMA-staged fused kernel lives in bbox_tma.cu; dyd_bbox_iou_fused
// picks between the two (DYD_FUSED=direct|tma, default tma).
#include <stdlib.h>
#include <string.h>

#include "kernels.cuh"
#include "k2_tile.cuh"

namespace dyd {

constexpr int CTA_THREADS = 256;
constexpr int CTA_WARPS = CTA_THREADS / 32;
constexpr int CROWD_SMEM_BOXES = 1024;

__device__ __forceinline__ void store_corner(double* pts, int64_t p, const Corner& c) {
    double2* o = reinterpret_cast<double2*>(pts + 4 * p);
    stg_stream_f64x2(o, make_double2(c.mnx, c.mny));
    stg_stream_f64x2(o + 1, make_double2(c.mxx, c.mxy));
}

// One G-lane group folds polygon [a, a+V) of the global vertex array (32/G register slots per
// lane).  Result in the group's lane 0.
template <bool ARG, int G>
__device__ __forceinline__ Corner fold_polygon(const double2* __restrict__ xy2, int64_t a, int V, int gl,
                                               unsigned gmask, CornerIdx& ci) {
    auto load = [&](int k) { return ldg_stream_f64x2(xy2 + a + k); };
    return group_bbox<G, 32 / G, ARG>(load, V, gl, gmask, ci);
}

// ------------------------------------------------------------------------------- K1
template <bool ARG, int G>
__global__ void __launch_bounds__(CTA_THREADS)
bbox_kernel(const int64_t* __restrict__ poly_off, const double2* __restrict__ xy2, int64_t n_poly,
            double* __restrict__ pts, uint8_t* __restrict__ valid, int32_t* __restrict__ arg) {
    const int lane = threadIdx.x & 31, gl = lane & (G - 1), g = lane / G;
    const unsigned gmask = (G == 32 ? FULL : ((1u << G) - 1u) << (g * G));
    const int64_t warp = (blockIdx.x * (int64_t)CTA_THREADS + threadIdx.x) >> 5;
    const int64_t p = warp * (32 / G) + g;
    if (p >= n_poly) return;                       // whole group leaves together
    const int64_t a = __ldg(poly_off + p), b = __ldg(poly_off + p + 1);
    const int64_t Vl = b - a;
    const int V = Vl > 0x7fffffff ? 0x7fffffff : (int)Vl;
    Corner c{0.0, 0.0, 0.0, 0.0};
    CornerIdx ci{-1, -1, -1, -1};
    if (V > 0) c = fold_polygon<ARG, G>(xy2, a, V, gl, gmask, ci);
    if (gl == 0) {
        store_corner(pts, p, c);
        valid[p] = V > 0 ? 1 : 0;
        if (ARG) *reinterpret_cast<int4*>(arg + 4 * p) = make_int4(ci.mnx, ci.mny, ci.mxx, ci.mxy);
    }
}

// Number of boxes before the first null bbox among objects [q0, q0+n) (warp-cooperative).
__device__ __forceinline__ int64_t valid_prefix(const uint8_t* __restrict__ valid, int64_t q0, int64_t n, int lane) {
    if (valid == nullptr) return n;
    for (int64_t base = 0; base < n; base += 32) {
        const int64_t j = base + lane;
        const bool bad = j < n && valid[q0 + j] == 0;
        const unsigned m = __ballot_sync(FULL, bad);
        if (m) return base + (__ffs(m) - 1);
    }
    return n;
}

__device__ __forceinline__ void defer_image(void* ws, int64_t img) {
    unsigned long long slot = atomicAdd(&reinterpret_cast<CrowdList*>(ws)->count, 1ULL);
    crowd_ids(ws)[slot] = (int)img;
}

// ------------------------------------------------------------------------------- K2 (warp per tile of images)
// A warp takes 32 consecutive images, packs them greedily into tiles of at most 32 boxes and
// K2_TM images (jump pointers, as in the fused kernel's pre-pass) and decides each tile with
// k2_tile_any: one lane per box, so small images do not leave most of the warp idle.  An image
// with more than 32 boxes goes to the block-per-image kernel.
constexpr int K2_TM = 8;
struct K2Warp {
    K2Tile t;
    long long off[33];                             // img_off slice of the warp's 32 images
    unsigned char nxt[32], start[33];
};
__global__ void __launch_bounds__(CTA_THREADS)
iou_tile_kernel(const int64_t* __restrict__ img_off, const double* __restrict__ pts, const uint8_t* __restrict__ valid,
                int64_t n_img, int64_t min_boxes, double thr, uint8_t* __restrict__ high,
                int32_t* __restrict__ count, void* ws) {
    __shared__ __align__(16) K2Warp sw[CTA_WARPS];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    K2Warp& s = sw[w];
    const bool zero_hits = 0.0 >= thr;
    const double2* pts2 = reinterpret_cast<const double2*>(pts);
    const int64_t n_seg = (n_img + 31) / 32;
    for (int64_t seg = blockIdx.x * (int64_t)CTA_WARPS + w; seg < n_seg; seg += (int64_t)gridDim.x * CTA_WARPS) {
        const int64_t i_begin = seg * 32;
        const int n_here = (int)min((int64_t)32, n_img - i_begin);
        {
            const int e = min(lane, n_here);
            s.off[e] = __ldg(img_off + i_begin + e);
            if (lane == 0) s.off[n_here] = __ldg(img_off + i_begin + n_here);
        }
        __syncwarp();
        if (lane < n_here) {                           // end of the maximal tile that starts at image `lane`
            const long long q0 = s.off[lane];
            int k = 1;
            while (k < K2_TM && lane + k < n_here && s.off[lane + k + 1] - q0 <= 32) ++k;
            s.nxt[lane] = (unsigned char)(lane + k);
        }
        __syncwarp();
        int cnt = 0;
        if (lane == 0) {
            for (int pos = 0; pos < n_here; pos = s.nxt[pos]) s.start[cnt++] = (unsigned char)pos;
            s.start[cnt] = (unsigned char)n_here;
        }
        cnt = __shfl_sync(FULL, cnt, 0);
        __syncwarp();
        for (int t = 0; t < cnt; ++t) {
            const int a = s.start[t], b = s.start[t + 1], ni = b - a;
            const int64_t q0 = s.off[a], np64 = s.off[b] - q0;
            if (np64 > 32) {                           // a single crowded image
                if (lane == 0) defer_image(ws, i_begin + a);
                continue;
            }
            const int np = (int)np64;
            bool nan_box = false, bad = false;
            if (lane < np) {
                const double2 p1 = ldg_stream_f64x2(pts2 + 2 * (q0 + lane)), p2 = ldg_stream_f64x2(pts2 + 2 * (q0 + lane) + 1);
                const Box bx = box_from_points(p1.x, p1.y, p2.x, p2.y);
                s.t.box_lo[lane] = make_double2(bx.x1, bx.y1); s.t.box_hi[lane] = make_double2(bx.x2, bx.y2);
                nan_box = (bx.x1 != bx.x1) | (bx.y1 != bx.y1) | (bx.x2 != bx.x2) | (bx.y2 != bx.y2);
                bad = valid != nullptr && valid[q0 + lane] == 0;
            }
            const unsigned inv = __ballot_sync(FULL, bad);
            const bool exact_pre = __any_sync(FULL, nan_box);
            int my_a = 0, my_n = 0, my_ne;
            if (lane < ni) {
                my_a = (int)(s.off[a + lane] - q0);
                my_n = (int)(s.off[a + lane + 1] - s.off[a + lane]);
            }
            const unsigned hits = k2_tile_any<K2_TM>(s.t, inv, exact_pre, my_a, my_n, ni, np, min_boxes, thr, zero_hits, lane, my_ne);
            if (lane < ni) { count[i_begin + a + lane] = my_ne; high[i_begin + a + lane] = (hits >> lane) & 1u; }
            __syncwarp();                              // tile scratch is rewritten by the next tile
        }
    }
}

// ------------------------------------------------------------------------------- K2 (block per crowded image)
// Processes the worklist: computes the valid prefix, writes count and high.
//
// Two forms.  Up to CROWD_SWEEP_MAX (512) boxes -- every image of config C4 -- the BINNED form below (crowd_sweep): only pairs
// that can overlap in x are met.  Larger images (up to 1024 boxes in shared memory) take the ALL-PAIRS form (crowd_tiled):
// box s meets s+1 .. s+(n-1)/2 (mod n); for even n the antipodal pair is taken by the lower half only -- every unordered pair
// exactly once, no pair table.  Register tiling: a thread owns 4 CONSECUTIVE boxes and walks the
// partners t = s0+1, s0+2, ... once; each partner (two 16-byte shared loads) is tested against all four own boxes, so the
// shared-memory traffic per pair is a quarter of the one-box-per-thread form and the loop is bound by the four fp64
// compares of the overlap pre-test (processor.py:329-333 reduces to them when no coordinate is NaN).  Boxes sit in shared
// memory de-interleaved by 4 (position (j & 3) * 256 + j / 4), which makes the lanes' stride-4 partner reads consecutive.
// Pairs that pass the pre-test (a fraction of a per cent) are queued and get the full IoU arithmetic from all threads after
// every chunk of 32 partners -- dense warps instead of one diverged lane -- and the block leaves as soon as one hits.
// (Tile widths of 2 and 3 for smaller images were measured and change nothing: idle lanes are filled by the other resident
// blocks.)  Images beyond shared memory, NaN coordinates and thr <= 0 take the generic loop at the end of the kernel.
constexpr int CROWD_THREADS = 128;
constexpr int CROWD_R = 4;
constexpr int CROWD_CHUNK = 32;
constexpr int CROWD_QCAP = 2048;
// all-pairs form: position (j % 4) * 256 + j / 4; binned form (R = 1): boxes in their own order
template <int R> __device__ __forceinline__ int crowd_pos_t(int j) { return (j % R) * (R == 4 ? CROWD_SMEM_BOXES / 4 : CROWD_THREADS) + j / R; }
__device__ __forceinline__ int crowd_pos(int j, int R) { return R == 4 ? crowd_pos_t<4>(j) : j; }
constexpr int CROWD_SWEEP_MAX = CROWD_SMEM_BOXES / 2;

// Binned form of the same question for images of at most CROWD_SWEEP_MAX boxes.  A pair can only reach a positive IoU if
// the boxes overlap in x, so the block buckets the boxes by x1 into CROWD_BINS bins with a counting sort in shared memory
// (bin = a monotone function of x1 between the image's smallest and largest x1), copies them in bin order into the upper
// half of the arrays, and every thread meets only the boxes behind its own up to the last bin its [x1, x2] touches.  Each
// unordered pair is taken from the side of the box that sits first in that array; the comparisons are the reference's own
// (processor.py:329-333), applied to a few per cent of the n(n-1)/2 pairs on crowd images.  Survivors of the pre-test are
// queued with their ORIGINAL box numbers and get the full IoU arithmetic from all threads, in the reference's argument order.
constexpr int CROWD_BINS = 64;
__device__ __forceinline__ int crowd_bin(double x, double mn, double scale) {
    double t = (x - mn) * scale;                              // monotone in x; scale == 0 puts every box into bin 0
    t = t < 0.0 ? 0.0 : t;
    t = t > (double)(CROWD_BINS - 1) ? (double)(CROWD_BINS - 1) : t;
    return (int)t;
}
// The register-tiled pre-test + queued exact test of one image whose boxes are in shared memory (layout of tile width R).
template <int R>
__device__ __forceinline__ void crowd_tiled(const double2* __restrict__ slo, const double2* __restrict__ shi, unsigned* __restrict__ queue,
                                            int& found, int& qn, int n, int half, bool even, double thr, int tid) {
    const int D = half + R;                              // last partner distance any own box can need
    for (int base = 0; base < n; base += CROWD_THREADS * R) {
        const int s0 = base + tid * R;
        double2 alo[R], ahi[R];
        int dmax[R];
#pragma unroll
        for (int k = 0; k < R; ++k) {
            const int row = s0 + k;
            if (row < n) { alo[k] = slo[crowd_pos_t<R>(row)]; ahi[k] = shi[crowd_pos_t<R>(row)]; dmax[k] = half + ((even && row < n / 2) ? 1 : 0); }
            else { alo[k] = make_double2(pos_inf(), pos_inf()); ahi[k] = make_double2(neg_inf(), neg_inf()); dmax[k] = 0; }   // meets nothing
        }
        const bool active = s0 < n;
        // one partner against the four own boxes; `rule` applies the pairing rule (needed at the two ends of the walk)
        auto meet = [&](int t, int d, bool rule) {
            const double2 blo = slo[crowd_pos_t<R>(t)], bhi = shi[crowd_pos_t<R>(t)];
            bool ov[R];
#pragma unroll
            for (int k = 0; k < R; ++k)
                ov[k] = ahi[k].x > blo.x && bhi.x > alo[k].x && ahi[k].y > blo.y && bhi.y > alo[k].y;
            bool some = false;
#pragma unroll
            for (int k = 0; k < R; ++k) some |= ov[k];
            if (some) {                                       // rare (a per cent of the partners): queue the survivors
                unsigned m = 0;
#pragma unroll
                for (int k = 0; k < R; ++k)
                    if (ov[k] && (!rule || (d - k >= 1 && d - k <= dmax[k]))) m |= 1u << k;
                if (m) {
                    unsigned slot = (unsigned)atomicAdd(&qn, __popc(m));
                    for (; m; m &= m - 1, ++slot) {
                        const int k = __ffs(m) - 1;
                        if (slot < (unsigned)CROWD_QCAP) queue[slot] = ((unsigned)(s0 + k) << 16) | (unsigned)t;
                        else if (iou_hits_cold(Box{slo[crowd_pos_t<R>(s0 + k)].x, slo[crowd_pos_t<R>(s0 + k)].y, shi[crowd_pos_t<R>(s0 + k)].x, shi[crowd_pos_t<R>(s0 + k)].y},
                                               Box{blo.x, blo.y, bhi.x, bhi.y}, thr, false)) found = 1;
                    }
                }
            }
        };
        for (int d0 = 1; d0 <= D; d0 += CROWD_CHUNK) {
            const int d1 = min(d0 + CROWD_CHUNK - 1, D);
            if (active) {
                // Pass 1, branch-free: walk the chunk's partners and note in a bit mask which of them overlap ANY own
                // box (pairing rule not applied yet).  Pass 2 revisits the noted partners -- about one in sixty --
                // applies the rule and queues the surviving pairs.  Keeping the rare work out of the walk keeps the
                // warps converged: the walk is 2 shared loads + 16 fp64 compares per partner.
                unsigned pend = 0;
                int t = s0 + d0;
                while (t >= n) t -= n;
                const int t_first = t;
#pragma unroll 4
                for (int d = d0; d <= d1; ++d) {
                    const double2 blo = slo[crowd_pos_t<R>(t)], bhi = shi[crowd_pos_t<R>(t)];
                    bool any = false;
#pragma unroll
                    for (int k = 0; k < R; ++k)
                        any |= ahi[k].x > blo.x && bhi.x > alo[k].x && ahi[k].y > blo.y && bhi.y > alo[k].y;
                    pend |= any ? (1u << (d - d0)) : 0u;
                    ++t; if (t >= n) t -= n;
                }
                for (; pend; pend &= pend - 1) {
                    const int b = __ffs(pend) - 1;
                    int tt = t_first + b;
                    while (tt >= n) tt -= n;
                    meet(tt, d0 + b, true);
                }
            }
            __syncthreads();
            const int nq = min(qn, CROWD_QCAP);
            for (int i = tid; i < nq; i += CROWD_THREADS) {
                const unsigned ent = queue[i];
                const int a = crowd_pos_t<R>((int)(ent >> 16)), b = crowd_pos_t<R>((int)(ent & 0xffffu));
                if (iou_hits(Box{slo[a].x, slo[a].y, shi[a].x, shi[a].y}, Box{slo[b].x, slo[b].y, shi[b].x, shi[b].y}, thr, false)) found = 1;
            }
            __syncthreads();
            const bool done = found != 0;
            if (tid == 0) qn = 0;
            __syncthreads();
            if (done) break;
        }
        if (found) break;                                      // uniform: read after the barrier above
    }
}

// The binned form as its own kernel: images of at most CROWD_SWEEP_MAX boxes.  A thread keeps its (at most four) boxes in
// registers from the global load to the scatter into the binned array, so shared memory holds ONE copy of the boxes (25 KB per
// block instead of the 42 KB of the kernel below, which keeps the 1024-box arrays of the all-pairs form): more resident blocks to
// hide the per-image latencies behind.  Larger images are left to iou_crowd_kernel, which skips the ones handled here.
constexpr int CROWD_OWN = CROWD_SWEEP_MAX / CROWD_THREADS;
#ifndef DYD_CROWD_MIN_BLOCKS
#define DYD_CROWD_MIN_BLOCKS 8
#endif
__global__ void __launch_bounds__(CROWD_THREADS, DYD_CROWD_MIN_BLOCKS)
iou_crowd_binned_kernel(const int64_t* __restrict__ img_off, const double* __restrict__ pts, const uint8_t* __restrict__ valid,
                        int64_t min_boxes, double thr, uint8_t* __restrict__ high, int32_t* __restrict__ count, void* ws) {
    __shared__ double2 qlo[CROWD_SWEEP_MAX], qhi[CROWD_SWEEP_MAX];        // (x1, y1) / (x2, y2) in bin order
    __shared__ __align__(16) unsigned queue[CROWD_QCAP];                   // first 192 words: bin counts / offsets / min-max exchange
    __shared__ unsigned short sidx[CROWD_SWEEP_MAX];                       // box number of every position
    __shared__ int found, has_nan, qn;
    __shared__ long long first_bad;
    const unsigned long long n_list = reinterpret_cast<CrowdList*>(ws)->count;
    const int* ids = crowd_ids(ws);
    const bool zero_hits = 0.0 >= thr;
    const int tid = threadIdx.x;
    int* cnt = reinterpret_cast<int*>(queue);
    int* off = cnt + CROWD_BINS;
    double* red = reinterpret_cast<double*>(cnt + 160);
    unsigned* q = queue + 192;
    constexpr int QCAP = CROWD_QCAP - 192;
    for (unsigned long long e = blockIdx.x; e < n_list; e += gridDim.x) {
        const int64_t img = ids[e];
        const int64_t q0 = img_off[img], n_all = img_off[img + 1] - q0;
        __syncthreads();                              // previous image fully consumed
        if (tid == 0) { found = 0; has_nan = 0; qn = 0; first_bad = n_all; }
        if (tid < CROWD_BINS) cnt[tid] = 0;
        __syncthreads();
        if (valid != nullptr) {
            long long mine = n_all;
            for (int64_t j = tid; j < n_all; j += CROWD_THREADS)
                if (valid[q0 + j] == 0) { mine = j; break; }
            if (mine < n_all) atomicMin(&first_bad, mine);
            __syncthreads();
        }
        const int64_t n64 = first_bad;
        if (n64 > CROWD_SWEEP_MAX) continue;          // the all-pairs kernel's image (uniform over the block)
        const int n = (int)n64;
        const bool want = n64 >= min_boxes && n >= 2;
        const double2* src = reinterpret_cast<const double2*>(pts + 4 * q0);
        if (want) {
            Box bx[CROWD_OWN];
            double mn = pos_inf(), mx = neg_inf();
            bool nan_here = false;
#pragma unroll
            for (int u = 0; u < CROWD_OWN; ++u) {
                const int j = tid + u * CROWD_THREADS;
                if (j < n) {
                    const double2 p1 = ldg_f64x2(src + 2 * j), p2 = ldg_f64x2(src + 2 * j + 1);
                    bx[u] = box_from_points(p1.x, p1.y, p2.x, p2.y);
                    nan_here |= bx[u].x1 != bx[u].x1 || bx[u].y1 != bx[u].y1 || bx[u].x2 != bx[u].x2 || bx[u].y2 != bx[u].y2;
                    mn = bx[u].x1 < mn ? bx[u].x1 : mn; mx = bx[u].x1 > mx ? bx[u].x1 : mx;
                }
            }
            if (nan_here) has_nan = 1;
            for (int o = 16; o > 0; o >>= 1) {
                const double a = __shfl_xor_sync(FULL, mn, o), b = __shfl_xor_sync(FULL, mx, o);
                mn = a < mn ? a : mn; mx = b > mx ? b : mx;
            }
            if ((tid & 31) == 0) { red[2 * (tid >> 5)] = mn; red[2 * (tid >> 5) + 1] = mx; }
            __syncthreads();
            if (!has_nan && !zero_hits) {
                mn = red[0]; mx = red[1];
                for (int w = 1; w < CROWD_THREADS / 32; ++w) { const double a = red[2 * w], b = red[2 * w + 1]; mn = a < mn ? a : mn; mx = b > mx ? b : mx; }
                double scale = (double)CROWD_BINS / (mx - mn);
                if (!(scale > 0.0) || !(scale < 1.0e300) || !(mn > -1.0e300)) scale = 0.0;      // one x1 value, or infinities: a single bin
                if (scale == 0.0) mn = 0.0;
                int mybin[CROWD_OWN], myrank[CROWD_OWN];
#pragma unroll
                for (int u = 0; u < CROWD_OWN; ++u) {
                    const int j = tid + u * CROWD_THREADS;
                    if (j < n) { mybin[u] = scale == 0.0 ? 0 : crowd_bin(bx[u].x1, mn, scale); myrank[u] = atomicAdd(&cnt[mybin[u]], 1); }
                }
                __syncthreads();
                if (tid < 32) {                                            // exclusive scan of the 64 counts by one warp
                    const int c0 = cnt[2 * tid], c1 = cnt[2 * tid + 1];
                    int x = c0 + c1;
                    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(FULL, x, o); if (tid >= o) x += y; }
                    off[2 * tid] = x - c0 - c1; off[2 * tid + 1] = x - c1;
                    if (tid == 31) off[CROWD_BINS] = x;
                }
                __syncthreads();
#pragma unroll
                for (int u = 0; u < CROWD_OWN; ++u) {
                    const int j = tid + u * CROWD_THREADS;
                    if (j < n) {
                        const int p = off[mybin[u]] + myrank[u];
                        qlo[p] = make_double2(bx[u].x1, bx[u].y1); qhi[p] = make_double2(bx[u].x2, bx[u].y2); sidx[p] = (unsigned short)j;
                    }
                }
                __syncthreads();
                for (int base = 0; base < n; base += CROWD_THREADS) {
                    const int i = base + tid;
                    if (i < n) {
                        const double2 alo = qlo[i], ahi = qhi[i];
                        const int b1 = scale == 0.0 ? 0 : crowd_bin(ahi.x, mn, scale);
                        const int p1 = off[b1 + 1];
                        // the boxes at positions i+1 .. p1-1 are all the partners box i has to meet (see crowd_sweep's comment)
                        for (int p = i + 1; p < p1; ++p) {
                            const double2 blo = qlo[p], bhi = qhi[p];
                            if (ahi.x > blo.x && bhi.x > alo.x && ahi.y > blo.y && bhi.y > alo.y) {
                                const bool i_first = sidx[i] < sidx[p];          // the reference meets the lower box number first
                                const unsigned lo = i_first ? (unsigned)i : (unsigned)p, hi = i_first ? (unsigned)p : (unsigned)i;
                                const unsigned slot = (unsigned)atomicAdd(&qn, 1);
                                if (slot < (unsigned)QCAP) q[slot] = (lo << 16) | hi;
                                else if (iou_hits_cold(Box{qlo[lo].x, qlo[lo].y, qhi[lo].x, qhi[lo].y}, Box{qlo[hi].x, qlo[hi].y, qhi[hi].x, qhi[hi].y}, thr, false)) found = 1;
                            }
                        }
                    }
                    __syncthreads();
                    const int nq = min(qn, QCAP);
                    for (int k2 = tid; k2 < nq; k2 += CROWD_THREADS) {
                        const unsigned ent = q[k2];
                        const int a = (int)(ent >> 16), b = (int)(ent & 0xffffu);
                        if (iou_hits(Box{qlo[a].x, qlo[a].y, qhi[a].x, qhi[a].y}, Box{qlo[b].x, qlo[b].y, qhi[b].x, qhi[b].y}, thr, false)) found = 1;
                    }
                    __syncthreads();
                    const bool done = found != 0;
                    if (tid == 0) qn = 0;
                    __syncthreads();
                    if (done) break;
                }
            } else {
                // NaN coordinates (exact selects in the reference's argument order) or thr <= 0: the generic loop on direct loads
                const int half = (n - 1) / 2;
                const bool even = (n & 1) == 0;
                auto load = [&](int j) {
                    const double2 p1 = ldg_f64x2(src + 2 * j), p2 = ldg_f64x2(src + 2 * j + 1);
                    return box_from_points(p1.x, p1.y, p2.x, p2.y);
                };
                for (int s2 = tid; s2 < n; s2 += CROWD_THREADS) {
                    const Box a = load(s2);
                    const int dmax = half + ((even && s2 < n / 2) ? 1 : 0);
                    bool mine = false;
                    int t = s2;
                    for (int d = 1; d <= dmax && !mine; ++d) {
                        ++t; if (t >= n) t -= n;
                        mine = iou_hits(a, load(t), thr, zero_hits);
                        if ((d & 31) == 0 && *(volatile int*)&found) break;
                    }
                    if (mine) found = 1;
                    if (*(volatile int*)&found) break;
                }
            }
        }
        __syncthreads();
        if (tid == 0) { high[img] = found ? 1 : 0; count[img] = n; }
    }
}

__global__ void __launch_bounds__(CROWD_THREADS)
iou_crowd_kernel(const int64_t* __restrict__ img_off, const double* __restrict__ pts, const uint8_t* __restrict__ valid,
                 int64_t min_boxes, double thr, uint8_t* __restrict__ high, int32_t* __restrict__ count, void* ws) {
    __shared__ double2 slo[CROWD_SMEM_BOXES], shi[CROWD_SMEM_BOXES];       // (x1, y1) / (x2, y2), de-interleaved by 4
    __shared__ __align__(16) unsigned queue[CROWD_QCAP];
    __shared__ int found, has_nan, qn;
    __shared__ long long first_bad;
    const unsigned long long n_list = reinterpret_cast<CrowdList*>(ws)->count;
    const int* ids = crowd_ids(ws);
    const bool zero_hits = 0.0 >= thr;
    const int tid = threadIdx.x;
    for (unsigned long long e = blockIdx.x; e < n_list; e += gridDim.x) {
        const int64_t img = ids[e];
        const int64_t q0 = img_off[img], n_all = img_off[img + 1] - q0;
        if (valid == nullptr && n_all <= CROWD_SWEEP_MAX) continue;    // cheap skip of the binned kernel's images
        __syncthreads();                              // previous image fully consumed
        if (tid == 0) { found = 0; has_nan = 0; qn = 0; first_bad = n_all; }
        __syncthreads();
        if (valid != nullptr) {
            long long mine = n_all;
            for (int64_t j = tid; j < n_all; j += CROWD_THREADS)
                if (valid[q0 + j] == 0) { mine = j; break; }
            if (mine < n_all) atomicMin(&first_bad, mine);
        }
        __syncthreads();
        const int64_t n64 = first_bad;
        const int n = n64 > 0x7fffffff ? 0x7fffffff : (int)n64;
        const bool want = n64 >= min_boxes && n >= 2;
        const double2* src = reinterpret_cast<const double2*>(pts + 4 * q0);
        const bool in_smem = n <= CROWD_SMEM_BOXES;
        // tile width: as many own boxes per thread as it takes to give every thread of the block work (2 .. 4); fewer own boxes
        // per thread means more shared-memory reads per pair but no idle lanes (n = 350: 117 busy threads with R = 3, 88 with 4)
        if (n64 <= CROWD_SWEEP_MAX) continue;         // iou_crowd_binned_kernel's image (uniform over the block)
        const int R = 4;                              // register-tiled all-pairs form, boxes de-interleaved by 4
        if (want && in_smem) {
            for (int j = tid; j < n; j += CROWD_THREADS) {
                double2 p1 = ldg_f64x2(src + 2 * j), p2 = ldg_f64x2(src + 2 * j + 1);
                Box bx = box_from_points(p1.x, p1.y, p2.x, p2.y);
                slo[crowd_pos(j, R)] = make_double2(bx.x1, bx.y1); shi[crowd_pos(j, R)] = make_double2(bx.x2, bx.y2);

                if (bx.x1 != bx.x1 || bx.y1 != bx.y1 || bx.x2 != bx.x2 || bx.y2 != bx.y2) has_nan = 1;
            }
        }
        __syncthreads();
        const int half = (n - 1) / 2;
        const bool even = (n & 1) == 0;
        if (want && in_smem && !has_nan && !zero_hits) {
            crowd_tiled<4>(slo, shi, queue, found, qn, n, half, even, thr, tid);
        } else if (want) {
            // ---- generic form: NaN coordinates (exact selects in the reference's argument order), thr <= 0, or an image
            //      too large for shared memory (direct loads)
            auto load = [&](int j) {
                if (in_smem) { const int pj = crowd_pos(j, R); return Box{slo[pj].x, slo[pj].y, shi[pj].x, shi[pj].y}; }
                double2 p1 = ldg_f64x2(src + 2 * j), p2 = ldg_f64x2(src + 2 * j + 1);
                return box_from_points(p1.x, p1.y, p2.x, p2.y);
            };
            for (int s = tid; s < n; s += CROWD_THREADS) {
                const Box a = load(s);
                const int dmax = half + ((even && s < n / 2) ? 1 : 0);
                bool mine = false;
                int t = s;
                for (int d = 1; d <= dmax && !mine; ++d) {
                    ++t; if (t >= n) t -= n;
                    mine = iou_hits(a, load(t), thr, zero_hits);
                    if ((d & 31) == 0 && *(volatile int*)&found) break;
                }
                if (mine) found = 1;
                if (*(volatile int*)&found) break;
            }
        }
        __syncthreads();
        if (tid == 0) { high[img] = found ? 1 : 0; count[img] = n; }
    }
}

// ------------------------------------------------------------------------------- fused K1 + K2 (direct loads)
template <bool ARG, int G>
__global__ void __launch_bounds__(CTA_THREADS)
fused_warp_kernel(const int64_t* __restrict__ img_off, const int64_t* __restrict__ poly_off,
                  const double2* __restrict__ xy2, int64_t n_img, int64_t min_boxes, double thr,
                  double* __restrict__ pts, uint8_t* __restrict__ valid, int32_t* __restrict__ arg,
                  uint8_t* __restrict__ high, int32_t* __restrict__ count, void* ws) {
    __shared__ __align__(16) double sbox[CTA_WARPS][WARP_BOX_CAP * 4];
    __shared__ __align__(16) unsigned short lut[PAIR_LUT_N + 8];
    load_pair_lut(lut, threadIdx.x, CTA_THREADS);
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int gl = lane & (G - 1), g = lane / G;
    const unsigned gmask = (G == 32 ? FULL : ((1u << G) - 1u) << (g * G));
    const bool zero_hits = 0.0 >= thr;
    for (int64_t img = blockIdx.x * (int64_t)CTA_WARPS + w; img < n_img; img += (int64_t)gridDim.x * CTA_WARPS) {
        const int64_t q0 = __ldg(img_off + img), q1 = __ldg(img_off + img + 1);
        const int64_t n = q1 - q0;
        int64_t n_eff = n;
        __syncwarp();
        for (int64_t base = 0; base < n; base += 32 / G) {
            const int64_t j = base + g;
            const bool active = j < n;
            int V = 1;
            if (active) {
                const int64_t p = q0 + j;
                const int64_t a = __ldg(poly_off + p), b = __ldg(poly_off + p + 1);
                const int64_t Vl = b - a;
                V = Vl > 0x7fffffff ? 0x7fffffff : (int)Vl;
                Corner c{0.0, 0.0, 0.0, 0.0};
                CornerIdx ci{-1, -1, -1, -1};
                if (V > 0) c = fold_polygon<ARG, G>(xy2, a, V, gl, gmask, ci);
                if (gl == 0) {
                    store_corner(pts, p, c);
                    valid[p] = V > 0 ? 1 : 0;
                    if (ARG) *reinterpret_cast<int4*>(arg + 4 * p) = make_int4(ci.mnx, ci.mny, ci.mxx, ci.mxy);
                    if (j < WARP_BOX_CAP) {
                        Box bx = box_from_points(c.mnx, c.mny, c.mxx, c.mxy);
                        double2* d = reinterpret_cast<double2*>(&sbox[w][4 * j]);
                        d[0] = make_double2(bx.x1, bx.y1); d[1] = make_double2(bx.x2, bx.y2);
                    }
                }
            }
            const unsigned bad = __ballot_sync(FULL, active && gl == 0 && V <= 0);
            if (bad && n_eff == n) n_eff = base + (__ffs(bad) - 1) / G;
        }
        __syncwarp();
        if (n > WARP_BOX_CAP) {                      // crowded: finished by the block kernel after this one
            if (lane == 0) defer_image(ws, img);
            continue;
        }
        bool hit = false;
        if (n_eff >= min_boxes && n_eff >= 2) hit = warp_any_pair(sbox[w], (int)n_eff, thr, zero_hits, lut, lane);
        if (lane == 0) { count[img] = (int32_t)n_eff; high[img] = hit ? 1 : 0; }
    }
}

// as many blocks as are resident at once (42 KB of shared memory per block: five per SM): the worklist is strided over the
// grid, so a block that only starts when another one has finished would double the time
static inline int crowd_grid() {
    static int per_sm = 0;
    if (per_sm == 0) {
        int v = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, iou_crowd_kernel, CROWD_THREADS, 0) != cudaSuccess || v < 1) v = 4;
        per_sm = v;
    }
    return NUM_SMS * per_sm;
}

static inline int crowd_binned_grid() {
    static int per_sm = 0;
    if (per_sm == 0) {
        int v = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, iou_crowd_binned_kernel, CROWD_THREADS, 0) != cudaSuccess || v < 1) v = 4;
        per_sm = v;
    }
    return NUM_SMS * per_sm;
}

static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

int launch_crowd(const int64_t* d_img_off, const double* d_pts, const uint8_t* d_valid, int64_t min_boxes,
                 double thr, uint8_t* d_high, int32_t* d_count, void* ws, cudaStream_t s) {
    iou_crowd_binned_kernel<<<crowd_binned_grid(), CROWD_THREADS, 0, s>>>(d_img_off, d_pts, d_valid, min_boxes, thr, d_high, d_count, ws);
    if (int rc = launch_check("iou_crowd_binned_kernel")) return rc;
    iou_crowd_kernel<<<crowd_grid(), CROWD_THREADS, 0, s>>>(d_img_off, d_pts, d_valid, min_boxes, thr, d_high, d_count, ws);
    return launch_check("iou_crowd_kernel");
}

template <bool ARG>
static int launch_bbox(int G, unsigned grid, cudaStream_t s, const int64_t* po, const double2* xy2, int64_t n,
                       double* pts, uint8_t* valid, int32_t* arg) {
    switch (G) {
        case 2: bbox_kernel<ARG, 2><<<grid, CTA_THREADS, 0, s>>>(po, xy2, n, pts, valid, arg); break;
        case 8: bbox_kernel<ARG, 8><<<grid, CTA_THREADS, 0, s>>>(po, xy2, n, pts, valid, arg); break;
        default: bbox_kernel<ARG, 4><<<grid, CTA_THREADS, 0, s>>>(po, xy2, n, pts, valid, arg); break;
    }
    return launch_check("bbox_kernel");
}

template <bool ARG>
static int launch_fused_direct(int G, unsigned grid, cudaStream_t s, const int64_t* io, const int64_t* po,
                               const double2* xy2, int64_t n_img, int64_t mb, double thr, double* pts, uint8_t* valid,
                               int32_t* arg, uint8_t* high, int32_t* count, void* ws) {
    switch (G) {
        case 2: fused_warp_kernel<ARG, 2><<<grid, CTA_THREADS, 0, s>>>(io, po, xy2, n_img, mb, thr, pts, valid, arg, high, count, ws); break;
        case 8: fused_warp_kernel<ARG, 8><<<grid, CTA_THREADS, 0, s>>>(io, po, xy2, n_img, mb, thr, pts, valid, arg, high, count, ws); break;
        default: fused_warp_kernel<ARG, 4><<<grid, CTA_THREADS, 0, s>>>(io, po, xy2, n_img, mb, thr, pts, valid, arg, high, count, ws); break;
    }
    return launch_check("fused_warp_kernel");
}

}  // namespace dyd

using namespace dyd;

extern "C" size_t dyd_iou_workspace_bytes(int64_t n_img) {
    if (n_img < 0) n_img = 0;
    return crowd_list_bytes(n_img) + tile_desc_bytes(n_img);
}

extern "C" int dyd_bbox_minmax(const int64_t* d_poly_off, const double* d_xy, int64_t n_poly,
                               double* d_pts, uint8_t* d_valid, int32_t* d_arg, void* stream) {
    DYD_REQUIRE(n_poly >= 0, DYD_E_ARG, "negative count");
    if (n_poly == 0) return 0;
    DYD_REQUIRE(d_poly_off && d_pts && d_valid, DYD_E_ARG, "null pointer");
    DYD_REQUIRE(((uintptr_t)d_xy & 15) == 0 && ((uintptr_t)d_pts & 15) == 0 && ((uintptr_t)d_arg & 15) == 0,
                DYD_E_ALIGN, "xy / pts / arg must be 16-byte aligned");
    const int G = env_int("DYD_GROUP", 4);
    const int per_warp = 32 / (G == 2 || G == 8 ? G : 4);
    const int64_t warps = (n_poly + per_warp - 1) / per_warp;
    const int64_t grid = (warps + CTA_WARPS - 1) / CTA_WARPS;
    DYD_REQUIRE(grid <= 0x7fffffff, DYD_E_ARG, "too many polygons for one launch");
    const double2* xy2 = reinterpret_cast<const double2*>(d_xy);
    if (d_arg) return launch_bbox<true>(G, (unsigned)grid, as_stream(stream), d_poly_off, xy2, n_poly, d_pts, d_valid, d_arg);
    return launch_bbox<false>(G, (unsigned)grid, as_stream(stream), d_poly_off, xy2, n_poly, d_pts, d_valid, nullptr);
}

extern "C" int dyd_iou_filter(const int64_t* d_img_off, const double* d_pts, const uint8_t* d_valid,
                              int64_t n_img, int64_t min_boxes, double thr, uint8_t* d_high, int32_t* d_count,
                              void* d_workspace, size_t workspace_bytes, void* stream) {
    DYD_REQUIRE(n_img >= 0 && n_img <= 0x7fffffff, DYD_E_ARG, "bad image count");
    if (n_img == 0) return 0;
    DYD_REQUIRE(d_img_off && d_high && d_count && d_workspace, DYD_E_ARG, "null pointer");
    DYD_REQUIRE(((uintptr_t)d_pts & 15) == 0 && ((uintptr_t)d_workspace & 15) == 0, DYD_E_ALIGN, "pts / workspace must be 16-byte aligned");
    DYD_REQUIRE(workspace_bytes >= dyd_iou_workspace_bytes(n_img), DYD_E_WORKSPACE, "workspace too small");
    cudaStream_t s = as_stream(stream);
    DYD_CUDA(cudaMemsetAsync(d_workspace, 0, sizeof(CrowdList), s));
    const int64_t want = ((n_img + 31) / 32 + CTA_WARPS - 1) / CTA_WARPS;
    const unsigned grid = (unsigned)(want < NUM_SMS * 16 ? want : NUM_SMS * 16);
    iou_tile_kernel<<<grid, CTA_THREADS, 0, s>>>(d_img_off, d_pts, d_valid, n_img, min_boxes, thr, d_high, d_count, d_workspace);
    if (int rc = launch_check("iou_tile_kernel")) return rc;
    return launch_crowd(d_img_off, d_pts, d_valid, min_boxes, thr, d_high, d_count, d_workspace, s);
}

extern "C" int dyd_bbox_iou_fused_ex(const int64_t* d_img_off, const int64_t* d_poly_off, const double* d_xy,
                                     int64_t n_img, int64_t n_poly, int64_t min_boxes, double thr,
                                     double* d_pts, uint8_t* d_valid, int32_t* d_arg, uint8_t* d_high, int32_t* d_count,
                                     void* d_workspace, size_t workspace_bytes, int32_t max_ctas, void* prepass_done_event, void* stream) {
    DYD_REQUIRE(n_img >= 0 && n_img <= 0x7fffffff && n_poly >= 0, DYD_E_ARG, "bad count");
    if (n_img == 0) return 0;
    DYD_REQUIRE(d_img_off && d_poly_off && d_pts && d_valid && d_high && d_count && d_workspace, DYD_E_ARG, "null pointer");
    DYD_REQUIRE(((uintptr_t)d_xy & 15) == 0 && ((uintptr_t)d_pts & 15) == 0 && ((uintptr_t)d_arg & 15) == 0 &&
                    ((uintptr_t)d_workspace & 15) == 0, DYD_E_ALIGN, "xy / pts / arg / workspace must be 16-byte aligned");
    DYD_REQUIRE(workspace_bytes >= dyd_iou_workspace_bytes(n_img), DYD_E_WORKSPACE, "workspace too small");
    cudaStream_t s = as_stream(stream);
    DYD_CUDA(cudaMemsetAsync(d_workspace, 0, sizeof(CrowdList), s));
    const char* variant = getenv("DYD_FUSED");
    const bool direct = variant && strcmp(variant, "direct") == 0;
    // the staged kernel bulk-copies offset slices: they must be 16-byte addressable
    const bool tma_ok = ((uintptr_t)d_img_off & 15) == 0 && ((uintptr_t)d_poly_off & 15) == 0;
    if (!direct && tma_ok) {
        if (int rc = launch_fused_tma(d_img_off, d_poly_off, d_xy, n_img, n_poly, min_boxes, thr, d_pts, d_valid, d_arg,
                                      d_high, d_count, d_workspace, max_ctas, reinterpret_cast<cudaEvent_t>(prepass_done_event), s)) return rc;
    } else {
        const int G = env_int("DYD_GROUP", 4);
        const int64_t want = (n_img + CTA_WARPS - 1) / CTA_WARPS;
        const unsigned grid = (unsigned)(want < NUM_SMS * 32 ? want : NUM_SMS * 32);
        const double2* xy2 = reinterpret_cast<const double2*>(d_xy);
        int rc = d_arg ? launch_fused_direct<true>(G, grid, s, d_img_off, d_poly_off, xy2, n_img, min_boxes, thr, d_pts, d_valid, d_arg, d_high, d_count, d_workspace)
                       : launch_fused_direct<false>(G, grid, s, d_img_off, d_poly_off, xy2, n_img, min_boxes, thr, d_pts, d_valid, nullptr, d_high, d_count, d_workspace);
        if (rc) return rc;
        if (prepass_done_event) DYD_CUDA(cudaEventRecord(reinterpret_cast<cudaEvent_t>(prepass_done_event), s));
    }
    return launch_crowd(d_img_off, d_pts, d_valid, min_boxes, thr, d_high, d_count, d_workspace, s);
}

extern "C" int dyd_fused_cta_times(uint64_t* h_times, int32_t n) {
    DYD_REQUIRE(h_times && n >= 0, DYD_E_ARG, "bad arguments");
    return fused_cta_times(reinterpret_cast<unsigned long long*>(h_times), n);
}

extern "C" int dyd_bbox_iou_fused(const int64_t* d_img_off, const int64_t* d_poly_off, const double* d_xy,
                                  int64_t n_img, int64_t n_poly, int64_t min_boxes, double thr,
                                  double* d_pts, uint8_t* d_valid, int32_t* d_arg, uint8_t* d_high, int32_t* d_count,
                                  void* d_workspace, size_t workspace_bytes, void* stream) {
    return dyd_bbox_iou_fused_ex(d_img_off, d_poly_off, d_xy, n_img, n_poly, min_boxes, thr, d_pts, d_valid, d_arg, d_high, d_count,
                                 d_workspace, workspace_bytes, 0, nullptr, stream);
}

extern "C" int dyd_fused_tile_modes(const void* d_workspace, int64_t n_img, uint64_t* d_counts3, void* stream) {
    DYD_REQUIRE(n_img >= 0 && d_workspace && d_counts3, DYD_E_ARG, "bad arguments");
    cudaStream_t s = as_stream(stream);
    DYD_CUDA(cudaMemsetAsync(d_counts3, 0, 3 * sizeof(uint64_t), s));
    return launch_tile_modes(d_workspace, n_img, reinterpret_cast<unsigned long long*>(d_counts3), s);
}

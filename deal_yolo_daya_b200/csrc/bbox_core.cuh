// Device building blocks shared by the K1 / K2 / fused kernels.
//
//   group_bbox_*   : an 8-lane group folds one polygon's vertices into its two corner
//                    points with CPython min()/max() semantics (processor.py:252-260)
//   iou_hits       : calculate_iou(...) >= thr (processor.py:328-339), fp64, no FMA,
//                    division replaced by a guarded product test that falls back to the
//                    exact IEEE division near the threshold
//   pair LUT       : unordered pairs (s < t) enumerated t-major, independent of n
#pragma once
#include "common.cuh"

namespace dyd {

constexpr int WARP_BOX_CAP = 64;         // images with more objects go to the block-per-image kernel
constexpr int PAIR_LUT_N = WARP_BOX_CAP * (WARP_BOX_CAP - 1) / 2;   // 2016
constexpr int IDX_NONE = 0x7fffffff;

__device__ __forceinline__ double pos_inf() { return __longlong_as_double(0x7ff0000000000000LL); }
__device__ __forceinline__ double neg_inf() { return __longlong_as_double(0xfff0000000000000LL); }
__device__ __forceinline__ bool is_neg_bits(double v) { return __double2hiint(v) < 0; }

struct Corner {          // (p1.x, p1.y, p2.x, p2.y) = (min_x, min_y, max_x, max_y)
    double mnx, mny, mxx, mxy;
};
struct CornerIdx {
    int mnx, mny, mxx, mxy;
};

// ---------------------------------------------------------------------------------------
// Fast path: V <= G*S, values only.  G lanes cooperate on one polygon; lane gl holds vertices
// gl, gl+G, gl+2G, ... (S slots), so each step of a group is one contiguous 16*G-byte request.
// CPython's fold equals
//     first vertex is NaN ? that NaN : min over the non-NaN values,
// so the lanes reduce with +-inf identities (a NaN never wins a strict comparison) and lane 0,
// which owns vertex 0, applies the NaN rule.  Equal values have equal bits except +-0.0; when a
// result is zero the group finds the FIRST zero in vertex order and takes its sign.
// The result is valid in the group's lane 0.
// ---------------------------------------------------------------------------------------
template <int G, int S, typename LoadFn>
__device__ __forceinline__ Corner group_bbox_fast(LoadFn load, int V, int gl, unsigned gmask) {
    double2 v[S];
    bool in[S];
#pragma unroll
    for (int t = 0; t < S; ++t) {
        int k = gl + G * t;
        in[t] = k < V;
        v[t] = in[t] ? load(k) : make_double2(0.0, 0.0);
    }
    Corner c{pos_inf(), pos_inf(), neg_inf(), neg_inf()};
#pragma unroll
    for (int t = 0; t < S; ++t) {
        if (in[t]) {
            c.mnx = v[t].x < c.mnx ? v[t].x : c.mnx;
            c.mxx = v[t].x > c.mxx ? v[t].x : c.mxx;
            c.mny = v[t].y < c.mny ? v[t].y : c.mny;
            c.mxy = v[t].y > c.mxy ? v[t].y : c.mxy;
        }
    }
#pragma unroll
    for (int off = G / 2; off >= 1; off >>= 1) {
        double o;
        o = __shfl_xor_sync(gmask, c.mnx, off); c.mnx = o < c.mnx ? o : c.mnx;
        o = __shfl_xor_sync(gmask, c.mxx, off); c.mxx = o > c.mxx ? o : c.mxx;
        o = __shfl_xor_sync(gmask, c.mny, off); c.mny = o < c.mny ? o : c.mny;
        o = __shfl_xor_sync(gmask, c.mxy, off); c.mxy = o > c.mxy ? o : c.mxy;
    }
    // signed zeros: every lane of the group sees value-equal results, so the vote is group-uniform
    bool zx = (c.mnx == 0.0) | (c.mxx == 0.0);
    bool zy = (c.mny == 0.0) | (c.mxy == 0.0);
    if (__any_sync(gmask, zx | zy)) {
        int kx = IDX_NONE, ky = IDX_NONE;          // (vertex index << 1) | sign of this lane's first zero
#pragma unroll
        for (int t = S - 1; t >= 0; --t) {
            int k = gl + G * t;
            if (in[t] && v[t].x == 0.0) kx = (k << 1) | (is_neg_bits(v[t].x) ? 1 : 0);
            if (in[t] && v[t].y == 0.0) ky = (k << 1) | (is_neg_bits(v[t].y) ? 1 : 0);
        }
#pragma unroll
        for (int off = G / 2; off >= 1; off >>= 1) {
            kx = min(kx, __shfl_xor_sync(gmask, kx, off));
            ky = min(ky, __shfl_xor_sync(gmask, ky, off));
        }
        double z0x = (kx & 1) ? -0.0 : 0.0, z0y = (ky & 1) ? -0.0 : 0.0;
        if (c.mnx == 0.0) c.mnx = z0x;
        if (c.mxx == 0.0) c.mxx = z0x;
        if (c.mny == 0.0) c.mny = z0y;
        if (c.mxy == 0.0) c.mxy = z0y;
    }
    if (gl == 0) {                                   // lane 0 holds vertex 0
        if (v[0].x != v[0].x) c.mnx = c.mxx = v[0].x;
        if (v[0].y != v[0].y) c.mny = c.mxy = v[0].y;
    }
    return c;
}

// ---------------------------------------------------------------------------------------
// General path: any V, tracks the vertex index of every extreme (needed for d_arg and for
// V > G*S).  (value, index) with "smaller index wins ties" is associative, so lanes fold their
// strided vertices and the group combines by butterfly; NaNs are skipped except at vertex 0.
// Result valid in the group's lane 0.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void take_min(double& m, int& im, double o, int io) {
    bool t = (io != IDX_NONE) && (im == IDX_NONE || o < m || (o == m && io < im));
    m = t ? o : m; im = t ? io : im;
}
__device__ __forceinline__ void take_max(double& m, int& im, double o, int io) {
    bool t = (io != IDX_NONE) && (im == IDX_NONE || o > m || (o == m && io < im));
    m = t ? o : m; im = t ? io : im;
}

template <int G, int S, typename LoadFn>
__device__ __forceinline__ Corner group_bbox_indexed(LoadFn load, int V, int gl, unsigned gmask, CornerIdx& ci) {
    Corner c{pos_inf(), pos_inf(), neg_inf(), neg_inf()};
    ci = CornerIdx{IDX_NONE, IDX_NONE, IDX_NONE, IDX_NONE};
    double2 v0 = make_double2(0.0, 0.0);
    for (int k0 = 0; k0 < V; k0 += S * G) {
        double2 v[S];
        bool in[S];
#pragma unroll
        for (int t = 0; t < S; ++t) {
            int k = k0 + gl + G * t;
            in[t] = k < V;
            v[t] = in[t] ? load(k) : make_double2(0.0, 0.0);
        }
        if (k0 == 0) v0 = v[0];
#pragma unroll
        for (int t = 0; t < S; ++t) {
            int k = k0 + gl + G * t;
            if (in[t]) {
                // first non-NaN value is taken unconditionally, later ones on strict comparison
                if (v[t].x == v[t].x) {
                    if (ci.mnx == IDX_NONE || v[t].x < c.mnx) { c.mnx = v[t].x; ci.mnx = k; }
                    if (ci.mxx == IDX_NONE || v[t].x > c.mxx) { c.mxx = v[t].x; ci.mxx = k; }
                }
                if (v[t].y == v[t].y) {
                    if (ci.mny == IDX_NONE || v[t].y < c.mny) { c.mny = v[t].y; ci.mny = k; }
                    if (ci.mxy == IDX_NONE || v[t].y > c.mxy) { c.mxy = v[t].y; ci.mxy = k; }
                }
            }
        }
    }
#pragma unroll
    for (int off = G / 2; off >= 1; off >>= 1) {
        double o; int io;
        o = __shfl_xor_sync(gmask, c.mnx, off); io = __shfl_xor_sync(gmask, ci.mnx, off); take_min(c.mnx, ci.mnx, o, io);
        o = __shfl_xor_sync(gmask, c.mxx, off); io = __shfl_xor_sync(gmask, ci.mxx, off); take_max(c.mxx, ci.mxx, o, io);
        o = __shfl_xor_sync(gmask, c.mny, off); io = __shfl_xor_sync(gmask, ci.mny, off); take_min(c.mny, ci.mny, o, io);
        o = __shfl_xor_sync(gmask, c.mxy, off); io = __shfl_xor_sync(gmask, ci.mxy, off); take_max(c.mxy, ci.mxy, o, io);
    }
    if (gl == 0) {
        if (v0.x != v0.x) { c.mnx = c.mxx = v0.x; ci.mnx = ci.mxx = 0; }
        if (v0.y != v0.y) { c.mny = c.mxy = v0.y; ci.mny = ci.mxy = 0; }
    }
    return c;
}

// One group folds a polygon of V vertices read through `load(k)`; picks the path.
template <int G, int S, bool ARG, typename LoadFn>
__device__ __forceinline__ Corner group_bbox(LoadFn load, int V, int gl, unsigned gmask, CornerIdx& ci) {
    if (ARG || V > G * S) return group_bbox_indexed<G, S>(load, V, gl, gmask, ci);
    return group_bbox_fast<G, S>(load, V, gl, gmask);
}

// ---------------------------------------------------------------------------------------
// IoU threshold test, bit-faithful to calculate_iou(...) >= thr.
// ---------------------------------------------------------------------------------------
struct Box {            // (x1, y1, x2, y2) after extract_boxes' min/max of the two points
    double x1, y1, x2, y2;
};

__device__ __forceinline__ Box box_from_points(double p1x, double p1y, double p2x, double p2y) {
    return Box{pymin(p1x, p2x), pymin(p1y, p2y), pymax(p1x, p2x), pymax(p1y, p2y)};
}

__device__ __forceinline__ bool iou_hits(const Box& a, const Box& b, double thr, bool zero_hits) {
    double xi1 = pymax(a.x1, b.x1), yi1 = pymax(a.y1, b.y1);
    double xi2 = pymin(a.x2, b.x2), yi2 = pymin(a.y2, b.y2);
    double dx = __dsub_rn(xi2, xi1), dy = __dsub_rn(yi2, yi1);
    double w = dx > 0.0 ? dx : 0.0, h = dy > 0.0 ? dy : 0.0;      // max(0, d)
    double inter = __dmul_rn(w, h);
    if (inter == 0.0) return zero_hits;                            // iou = 0.0
    double area1 = __dmul_rn(__dsub_rn(a.x2, a.x1), __dsub_rn(a.y2, a.y1));
    double area2 = __dmul_rn(__dsub_rn(b.x2, b.x1), __dsub_rn(b.y2, b.y1));
    double uni = __dsub_rn(__dadd_rn(area1, area2), inter);
    if (uni == 0.0) return zero_hits;
    // Guarded shortcut: with p = RN(thr*uni), inter outside p*(1 -+ 2^-50) decides the comparison
    // RN(inter/uni) >= thr without dividing (every product below is in the normal range, so each
    // rounding is within 2^-53 relative and RN is monotonic); anything closer to the threshold,
    // or outside the checked range, takes the exact IEEE division.
    const double lo = 1e-280, hi = 1e280;
    double p = __dmul_rn(thr, uni);
    if (p > lo && p < hi && inter > lo && inter < hi) {
        if (inter > __dmul_rn(p, 1.0 + 0x1p-50)) return true;
        if (inter < __dmul_rn(p, 1.0 - 0x1p-50)) return false;
    }
    return __ddiv_rn(inter, uni) >= thr;
}

// Pair k -> (s, t), s < t, ordered by t then s: k = t(t-1)/2 + s.  Valid for an image with n
// boxes iff t < n, i.e. k < n(n-1)/2; the table does not depend on n.  Built at compile time,
// lives in global memory (4 KB, L2 resident) and is copied to shared memory by each CTA.
struct PairTable {
    unsigned short st[PAIR_LUT_N + 8];     // (t << 8) | s ; padded to a multiple of 16 bytes
};
constexpr PairTable make_pair_table() {
    PairTable t{};
    int k = 0;
    for (int tt = 1; tt < WARP_BOX_CAP; ++tt)
        for (int s = 0; s < tt; ++s) t.st[k++] = (unsigned short)((tt << 8) | s);
    return t;
}
__device__ const PairTable g_pair_table = make_pair_table();
static_assert(sizeof(PairTable) % 16 == 0, "pair table must be copyable in 16-byte pieces");

__device__ __forceinline__ void load_pair_lut(unsigned short* st, int tid, int nthreads) {
    const uint4* src = reinterpret_cast<const uint4*>(g_pair_table.st);
    uint4* dst = reinterpret_cast<uint4*>(st);
    for (int k = tid; k < (int)(sizeof(PairTable) / 16); k += nthreads) dst[k] = __ldg(src + k);
}

// Warp-level any-pair test over n boxes stored as rows of 4 doubles in shared memory.
__device__ __forceinline__ bool warp_any_pair(const double* sb, int n, double thr, bool zero_hits,
                                              const unsigned short* lut, int lane) {
    int npairs = n * (n - 1) / 2;
    bool hit = false;
    for (int k0 = 0; k0 < npairs; k0 += 32) {
        int k = k0 + lane;
        bool mine = false;
        if (k < npairs) {
            unsigned st = lut[k];
            int s = st & 0xff, t = st >> 8;
            const double2* ps = reinterpret_cast<const double2*>(sb + 4 * s);
            const double2* pt = reinterpret_cast<const double2*>(sb + 4 * t);
            double2 s0 = ps[0], s1 = ps[1], t0 = pt[0], t1 = pt[1];
            mine = iou_hits(Box{s0.x, s0.y, s1.x, s1.y}, Box{t0.x, t0.y, t1.x, t1.y}, thr, zero_hits);
        }
        if (__any_sync(FULL, mine)) { hit = true; break; }
    }
    return hit;
}

}  // namespace dyd

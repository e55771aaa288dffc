// Device twin of deal_yolo_daya_b200/synth.py: seeded synthetic annotation tables generated
// straight into HBM.  Integer mixing + separately rounded IEEE double operations
// only (the library is compiled with -fmad=false), so the output is bit-identical to the numpy
// generator and any slice can be re-created on the host for parity checks.
#include "common.cuh"
#include "../../include/dyd_synth.h"

namespace dyd {

constexpr int SY_THREADS = 256;
constexpr double IMG_W = 1920.0, IMG_H = 1080.0;
constexpr int N_LABELS = 80, MAX_POLYS = 50, MIN_VERTS = 4, VERT_SPAN = 29;
enum : uint64_t {
    TAG_NPOLY = 0x0A, TAG_DUPBOX = 0x0B, TAG_NVERT = 0x0C, TAG_CX = 0x0D, TAG_CY = 0x0E, TAG_R = 0x0F,
    TAG_LABEL = 0x10, TAG_URLDUP = 0x11, TAG_URLPICK = 0x12, TAG_REFHIT = 0x13, TAG_REFPICK = 0x14,
    TAG_NBOX_CROWD = 0x15, TAG_VERT = 0x1000, TAG_JITTER = 0x2000
};

__device__ __forceinline__ uint64_t rnd(uint64_t seed, uint64_t a, uint64_t tag, uint64_t k) {
    return mix64(mix64(mix64(seed + a) + tag) + k);
}
__device__ __forceinline__ double u01(uint64_t h) { return (double)(h >> 11) * 0x1p-53; }
__device__ __forceinline__ double clip(double v, double hi) { v = v < 0.0 ? 0.0 : v; return v > hi ? hi : v; }

__device__ __forceinline__ int n_polys_of(uint64_t seed, uint64_t id, const uint64_t* __restrict__ thr, int n_thr) {
    const uint64_t h = rnd(seed, id, TAG_NPOLY, 0) >> 11;
    int n = 0;
    for (int i = 0; i < n_thr; ++i) n += thr[i] <= h ? 1 : 0;     // searchsorted(side="right")
    return n < 1 ? 1 : (n > MAX_POLYS ? MAX_POLYS : n);
}
__device__ __forceinline__ bool dupbox_of(uint64_t seed, uint64_t id, int64_t n) {
    return u01(rnd(seed, id, TAG_DUPBOX, 0)) < 0.10 && n >= 2;
}

__global__ void __launch_bounds__(SY_THREADS)
synth_counts_kernel(uint64_t seed, int64_t first_img, int64_t n_img, const uint64_t* __restrict__ thr, int n_thr,
                    int64_t* __restrict__ npoly) {
    const int64_t i = blockIdx.x * (int64_t)SY_THREADS + threadIdx.x;
    if (i < n_img) npoly[i] = n_polys_of(seed, (uint64_t)(first_img + i), thr, n_thr);
}

__global__ void __launch_bounds__(SY_THREADS)
synth_nvert_kernel(uint64_t seed, int64_t first_img, int64_t n_img, const int64_t* __restrict__ img_off,
                   int64_t* __restrict__ nvert) {
    const int64_t i = blockIdx.x * (int64_t)SY_THREADS + threadIdx.x;
    if (i >= n_img) return;
    const uint64_t id = (uint64_t)(first_img + i);
    const int64_t q0 = img_off[i], n = img_off[i + 1] - q0;
    const bool dup = dupbox_of(seed, id, n);
    for (int64_t j = 0; j < n; ++j) {
        const uint64_t src = (dup && j == n - 1) ? 0 : (uint64_t)j;
        nvert[q0 + j] = MIN_VERTS + (int64_t)(rnd(seed, id, TAG_NVERT, src) % VERT_SPAN);
    }
}

// one warp per image
__global__ void __launch_bounds__(SY_THREADS)
synth_fill_kernel(uint64_t seed, int64_t first_img, int64_t n_img, const int64_t* __restrict__ img_off,
                  const int64_t* __restrict__ poly_off, double2* __restrict__ xy2, int32_t* __restrict__ label_id) {
    const int lane = threadIdx.x & 31;
    const int64_t i = (blockIdx.x * (int64_t)SY_THREADS + threadIdx.x) >> 5;
    if (i >= n_img) return;
    const uint64_t id = (uint64_t)(first_img + i);
    const int64_t q0 = img_off[i], n = img_off[i + 1] - q0;
    const bool dup = dupbox_of(seed, id, n);
    for (int64_t j = 0; j < n; ++j) {
        const bool copy = dup && j == n - 1;
        const uint64_t src = copy ? 0 : (uint64_t)j;
        const double cx = u01(rnd(seed, id, TAG_CX, src)) * IMG_W;
        const double cy = u01(rnd(seed, id, TAG_CY, src)) * IMG_H;
        const double r = 5.0 + u01(rnd(seed, id, TAG_R, src)) * 195.0;
        const int64_t a = poly_off[q0 + j], V = poly_off[q0 + j + 1] - a;
        if (lane == 0) label_id[q0 + j] = (int32_t)(rnd(seed, id, TAG_LABEL, (uint64_t)j) % N_LABELS);
        for (int64_t k = lane; k < V; k += 32) {
            const double ux = u01(rnd(seed, id, TAG_VERT + src, 2 * (uint64_t)k));
            const double uy = u01(rnd(seed, id, TAG_VERT + src, 2 * (uint64_t)k + 1));
            double x = clip(cx + (2.0 * ux - 1.0) * r, IMG_W);
            double y = clip(cy + (2.0 * uy - 1.0) * r, IMG_H);
            if (copy) {
                const double jx = u01(rnd(seed, id, TAG_JITTER, 2 * (uint64_t)k));
                const double jy = u01(rnd(seed, id, TAG_JITTER, 2 * (uint64_t)k + 1));
                const double amp = 0.005 * r;
                x = clip(x + (2.0 * jx - 1.0) * amp, IMG_W);
                y = clip(y + (2.0 * jy - 1.0) * amp, IMG_H);
            }
            xy2[a + k] = make_double2(x, y);
        }
    }
}

__device__ __forceinline__ int n_digits(uint64_t v) { int d = 1; while (v >= 10) { v /= 10; ++d; } return d; }

__global__ void __launch_bounds__(SY_THREADS)
synth_urls_kernel(uint64_t seed, int64_t first_row, int64_t n, int64_t n_main_for_ref,
                  int64_t* __restrict__ url_id, int64_t* __restrict__ len) {
    const int64_t r = blockIdx.x * (int64_t)SY_THREADS + threadIdx.x;
    if (r >= n) return;
    const uint64_t id = (uint64_t)(first_row + r);
    uint64_t out;
    if (n_main_for_ref < 0) {
        const bool isdup = u01(rnd(seed, id, TAG_URLDUP, 0)) < 0.05 && id > 0;
        const uint64_t pick = rnd(seed, id, TAG_URLPICK, 0) % (id > 0 ? id : 1);
        out = isdup ? pick : id;
    } else {
        const bool hit = u01(rnd(seed, id, TAG_REFHIT, 0)) < 0.10;
        const uint64_t nm = (uint64_t)(n_main_for_ref > 0 ? n_main_for_ref : 1);
        const uint64_t pick = rnd(seed, id, TAG_REFPICK, 0) % nm;
        out = hit ? pick : (uint64_t)n_main_for_ref + id;
    }
    url_id[r] = (int64_t)out;
    if (len) len[r] = 24 + n_digits(out) + 4;
}

__global__ void __launch_bounds__(SY_THREADS)
synth_url_bytes_kernel(const int64_t* __restrict__ url_id, const int64_t* __restrict__ off, int64_t n,
                       uint8_t* __restrict__ bytes) {
    const int64_t r = blockIdx.x * (int64_t)SY_THREADS + threadIdx.x;
    if (r >= n) return;
    const char head[] = "https://img.example.com/";
    const char tail[] = ".jpg";
    uint8_t* p = bytes + off[r];
    for (int i = 0; i < 24; ++i) p[i] = (uint8_t)head[i];
    uint64_t v = (uint64_t)url_id[r];
    const int d = n_digits(v);
    for (int i = d - 1; i >= 0; --i) { p[24 + i] = (uint8_t)('0' + v % 10); v /= 10; }
    for (int i = 0; i < 4; ++i) p[24 + d + i] = (uint8_t)tail[i];
}

__global__ void __launch_bounds__(SY_THREADS)
synth_crowd_kernel(uint64_t seed, int64_t first_img, int64_t n_img, int lo, int hi,
                   const int64_t* __restrict__ img_off, int64_t* __restrict__ nbox, double* __restrict__ pts) {
    const int lane = threadIdx.x & 31;
    const int64_t i = (blockIdx.x * (int64_t)SY_THREADS + threadIdx.x) >> 5;
    if (i >= n_img) return;
    const uint64_t id = (uint64_t)(first_img + i);
    const int64_t n = lo + (int64_t)(rnd(seed, id, TAG_NBOX_CROWD, 0) % (uint64_t)(hi - lo + 1));
    if (img_off == nullptr) { if (lane == 0) nbox[i] = n; return; }
    const int64_t q0 = img_off[i];
    for (int64_t j = lane; j < n; j += 32) {
        const double cx = u01(rnd(seed, id, TAG_CX, (uint64_t)j)) * IMG_W;
        const double cy = u01(rnd(seed, id, TAG_CY, (uint64_t)j)) * IMG_H;
        const double hw = 4.0 + u01(rnd(seed, id, TAG_R, (uint64_t)j)) * 36.0;
        const double hh = 4.0 + u01(rnd(seed, id, TAG_NVERT, (uint64_t)j)) * 36.0;
        double x1 = cx - hw, y1 = cy - hh, x2 = cx + hw, y2 = cy + hh;
        x1 = x1 < 0.0 ? 0.0 : x1; y1 = y1 < 0.0 ? 0.0 : y1;
        x2 = x2 > IMG_W ? IMG_W : x2; y2 = y2 > IMG_H ? IMG_H : y2;
        double* o = pts + 4 * (q0 + j);
        o[0] = x1; o[1] = y1; o[2] = x2; o[3] = y2;
    }
}

static inline unsigned blocks(int64_t n, int per_block) { return (unsigned)((n + per_block - 1) / per_block); }

}  // namespace dyd

using namespace dyd;

extern "C" int dyd_synth_counts(uint64_t seed, int64_t first_img, int64_t n_img, const uint64_t* d_pois_thr,
                                int32_t n_thr, int64_t* d_npoly, void* stream) {
    DYD_REQUIRE(n_img >= 0 && n_thr >= 0, DYD_E_ARG, "negative count");
    if (n_img == 0) return 0;
    DYD_REQUIRE(d_pois_thr && d_npoly, DYD_E_ARG, "null pointer");
    synth_counts_kernel<<<blocks(n_img, SY_THREADS), SY_THREADS, 0, as_stream(stream)>>>(seed, first_img, n_img, d_pois_thr, n_thr, d_npoly);
    return launch_check("synth_counts_kernel");
}

extern "C" int dyd_synth_nvert(uint64_t seed, int64_t first_img, int64_t n_img, const int64_t* d_img_off,
                               int64_t* d_nvert, void* stream) {
    DYD_REQUIRE(n_img >= 0, DYD_E_ARG, "negative count");
    if (n_img == 0) return 0;
    DYD_REQUIRE(d_img_off && d_nvert, DYD_E_ARG, "null pointer");
    synth_nvert_kernel<<<blocks(n_img, SY_THREADS), SY_THREADS, 0, as_stream(stream)>>>(seed, first_img, n_img, d_img_off, d_nvert);
    return launch_check("synth_nvert_kernel");
}

extern "C" int dyd_synth_fill(uint64_t seed, int64_t first_img, int64_t n_img, const int64_t* d_img_off,
                              const int64_t* d_poly_off, double* d_xy, int32_t* d_label_id, void* stream) {
    DYD_REQUIRE(n_img >= 0, DYD_E_ARG, "negative count");
    if (n_img == 0) return 0;
    DYD_REQUIRE(d_img_off && d_poly_off && d_xy && d_label_id, DYD_E_ARG, "null pointer");
    DYD_REQUIRE(((uintptr_t)d_xy & 15) == 0, DYD_E_ALIGN, "xy must be 16-byte aligned");
    synth_fill_kernel<<<blocks(n_img, SY_THREADS / 32), SY_THREADS, 0, as_stream(stream)>>>(
        seed, first_img, n_img, d_img_off, d_poly_off, reinterpret_cast<double2*>(d_xy), d_label_id);
    return launch_check("synth_fill_kernel");
}

extern "C" int dyd_synth_urls(uint64_t seed, int64_t first_row, int64_t n, int64_t n_main_for_ref,
                              int64_t* d_url_id, int64_t* d_len, void* stream) {
    DYD_REQUIRE(n >= 0, DYD_E_ARG, "negative count");
    if (n == 0) return 0;
    DYD_REQUIRE(d_url_id, DYD_E_ARG, "null pointer");
    synth_urls_kernel<<<blocks(n, SY_THREADS), SY_THREADS, 0, as_stream(stream)>>>(seed, first_row, n, n_main_for_ref, d_url_id, d_len);
    return launch_check("synth_urls_kernel");
}

extern "C" int dyd_synth_url_bytes(const int64_t* d_url_id, const int64_t* d_off, int64_t n, uint8_t* d_bytes, void* stream) {
    DYD_REQUIRE(n >= 0, DYD_E_ARG, "negative count");
    if (n == 0) return 0;
    DYD_REQUIRE(d_url_id && d_off && d_bytes, DYD_E_ARG, "null pointer");
    synth_url_bytes_kernel<<<blocks(n, SY_THREADS), SY_THREADS, 0, as_stream(stream)>>>(d_url_id, d_off, n, d_bytes);
    return launch_check("synth_url_bytes_kernel");
}

extern "C" int dyd_synth_crowd(uint64_t seed, int64_t first_img, int64_t n_img, int32_t lo, int32_t hi,
                               const int64_t* d_img_off, int64_t* d_nbox, double* d_pts, void* stream) {
    DYD_REQUIRE(n_img >= 0 && lo >= 0 && hi >= lo, DYD_E_ARG, "bad arguments");
    if (n_img == 0) return 0;
    DYD_REQUIRE((d_img_off == nullptr && d_nbox) || (d_img_off && d_pts), DYD_E_ARG, "null pointer");
    synth_crowd_kernel<<<blocks(n_img, SY_THREADS / 32), SY_THREADS, 0, as_stream(stream)>>>(seed, first_img, n_img, lo, hi, d_img_off, d_nbox, d_pts);
    return launch_check("synth_crowd_kernel");
}

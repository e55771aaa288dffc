// Host-buffer entry points: the calls the Python drop-in makes for tables that live in host
// memory.  The image range is cut into chunks that are pushed through H2D -> fused K1+K2 ->
// D2H on three internal streams, so the copy engines and the SMs work on different chunks at
// the same time.  Transient device buffers come from a library-owned CUDA memory pool per device
// (stream-ordered allocation, freed before returning).  The pool keeps the freed blocks cached
// between calls -- re-creating gigabytes of physical memory per call cost 0.1-1 s on the B200 box --
// and dyd_host_release() hands them back to the driver.
#include <algorithm>
#include <mutex>
#include <vector>

#include "kernels.cuh"

namespace dyd {

constexpr int NSLOT = 3;
constexpr int MAX_DEVICES = 64;

static std::mutex g_pool_mutex;
static cudaMemPool_t g_pools[MAX_DEVICES] = {};

// The calling thread's current device's pool (created on first use; cached blocks are never trimmed
// implicitly: release threshold = UINT64_MAX).
static int host_pool(cudaMemPool_t* out) {
    int dev = 0;
    DYD_CUDA(cudaGetDevice(&dev));
    DYD_REQUIRE(dev >= 0 && dev < MAX_DEVICES, DYD_E_ARG, "device index out of range");
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    if (!g_pools[dev]) {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        DYD_CUDA(cudaMemPoolCreate(&g_pools[dev], &props));
        unsigned long long keep = ~0ULL;
        DYD_CUDA(cudaMemPoolSetAttribute(g_pools[dev], cudaMemPoolAttrReleaseThreshold, &keep));
    }
    *out = g_pools[dev];
    return 0;
}

struct Slot3 {
    cudaStream_t stream = nullptr;
    int64_t* img_off = nullptr;
    int64_t* poly_off = nullptr;
    double* xy = nullptr;
    double* pts = nullptr;
    uint8_t* valid = nullptr;
    int32_t* arg = nullptr;
    uint8_t* high = nullptr;
    int32_t* count = nullptr;
    void* ws = nullptr;
};

struct SlotGuard {
    Slot3 s[NSLOT];
    ~SlotGuard() {
        for (auto& x : s) {
            if (!x.stream) continue;
            void* bufs[] = {x.img_off, x.poly_off, x.xy, x.pts, x.valid, x.arg, x.high, x.count, x.ws};
            for (void* b : bufs) if (b) cudaFreeAsync(b, x.stream);
            cudaStreamSynchronize(x.stream);
            cudaStreamDestroy(x.stream);
        }
    }
};

}  // namespace dyd

using namespace dyd;

extern "C" int dyd_bbox_iou_host_ex(const int64_t* h_img_off, const int64_t* h_poly_off, const double* h_xy,
                                    int64_t n_img, int64_t min_boxes, double thr,
                                    double* h_pts, uint8_t* h_valid, int32_t* h_arg,
                                    uint8_t* h_high, int32_t* h_count, int64_t chunk_images, int64_t* h_tile_modes) {
    DYD_REQUIRE(n_img >= 0, DYD_E_ARG, "negative count");
    if (h_tile_modes) h_tile_modes[0] = h_tile_modes[1] = h_tile_modes[2] = 0;
    if (n_img == 0) return 0;
    DYD_REQUIRE(h_img_off && h_poly_off && h_high && h_count, DYD_E_ARG, "null pointer");
    if (chunk_images <= 0) chunk_images = 16384;
    const int64_t n_chunks = (n_img + chunk_images - 1) / chunk_images;
    int64_t max_img = 0, max_poly = 0, max_vert = 0;
    for (int64_t c = 0; c < n_chunks; ++c) {
        const int64_t i0 = c * chunk_images, i1 = std::min(n_img, i0 + chunk_images);
        const int64_t q0 = h_img_off[i0], q1 = h_img_off[i1];
        DYD_REQUIRE(q1 >= q0, DYD_E_ARG, "img_off not monotone");
        const int64_t v0 = h_poly_off[q0], v1 = h_poly_off[q1];
        DYD_REQUIRE(v1 >= v0, DYD_E_ARG, "poly_off not monotone");
        max_img = std::max(max_img, i1 - i0); max_poly = std::max(max_poly, q1 - q0); max_vert = std::max(max_vert, v1 - v0);
    }
    DYD_REQUIRE(max_vert == 0 || h_xy, DYD_E_ARG, "null pointer");
    cudaMemPool_t pool;
    if (int rc = host_pool(&pool)) return rc;
    SlotGuard guard;
    const size_t ws_bytes = dyd_iou_workspace_bytes(max_img);
    const int nslot = (int)std::min<int64_t>(NSLOT, n_chunks);
    for (int k = 0; k < nslot; ++k) {
        Slot3& s = guard.s[k];
        DYD_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        DYD_CUDA(cudaMallocFromPoolAsync((void**)&s.img_off, sizeof(int64_t) * (max_img + 1), pool, s.stream));
        // one spare entry in front: a chunk whose first object number is odd is stored from entry 1 on, so that the
        // rebased array (indexed with global object numbers) stays 16-byte aligned for the bulk copies
        DYD_CUDA(cudaMallocFromPoolAsync((void**)&s.poly_off, sizeof(int64_t) * (max_poly + 2), pool, s.stream));
        DYD_CUDA(cudaMallocFromPoolAsync((void**)&s.xy, sizeof(double) * 2 * std::max<int64_t>(max_vert, 1), pool, s.stream));
        DYD_CUDA(cudaMallocFromPoolAsync((void**)&s.pts, sizeof(double) * 4 * std::max<int64_t>(max_poly, 1), pool, s.stream));
        DYD_CUDA(cudaMallocFromPoolAsync((void**)&s.valid, std::max<int64_t>(max_poly, 1), pool, s.stream));
        if (h_arg) DYD_CUDA(cudaMallocFromPoolAsync((void**)&s.arg, sizeof(int32_t) * 4 * std::max<int64_t>(max_poly, 1), pool, s.stream));
        DYD_CUDA(cudaMallocFromPoolAsync((void**)&s.high, max_img, pool, s.stream));
        DYD_CUDA(cudaMallocFromPoolAsync((void**)&s.count, sizeof(int32_t) * max_img, pool, s.stream));
        DYD_CUDA(cudaMallocFromPoolAsync(&s.ws, ws_bytes, pool, s.stream));
    }
    unsigned long long* d_modes = nullptr;
    if (h_tile_modes) {
        DYD_CUDA(cudaMallocFromPoolAsync((void**)&d_modes, 3 * sizeof(unsigned long long), pool, guard.s[0].stream));
        DYD_CUDA(cudaMemsetAsync(d_modes, 0, 3 * sizeof(unsigned long long), guard.s[0].stream));
        DYD_CUDA(cudaStreamSynchronize(guard.s[0].stream));
    }
    for (int64_t c = 0; c < n_chunks; ++c) {
        Slot3& s = guard.s[c % nslot];
        const int64_t i0 = c * chunk_images, i1 = std::min(n_img, i0 + chunk_images), ni = i1 - i0;
        const int64_t q0 = h_img_off[i0], q1 = h_img_off[i1], nq = q1 - q0;
        const int64_t v0 = h_poly_off[q0], v1 = h_poly_off[q1], nv = v1 - v0;
        int64_t* const d_poly = s.poly_off + (q0 & 1);
        DYD_CUDA(cudaMemcpyAsync(s.img_off, h_img_off + i0, sizeof(int64_t) * (ni + 1), cudaMemcpyHostToDevice, s.stream));
        DYD_CUDA(cudaMemcpyAsync(d_poly, h_poly_off + q0, sizeof(int64_t) * (nq + 1), cudaMemcpyHostToDevice, s.stream));
        if (nv) DYD_CUDA(cudaMemcpyAsync(s.xy, h_xy + 2 * v0, sizeof(double) * 2 * nv, cudaMemcpyHostToDevice, s.stream));
        // The kernels index with the table's global object / vertex numbers: rebase the chunk buffers.  The object
        // count handed over is the END of the rebased poly_off array (q1 entries + 1 are addressable), which is what
        // the staged kernel's "bulk copy stays inside the array" test compares global indices with.
        int rc = dyd_bbox_iou_fused(s.img_off, d_poly - q0, s.xy - 2 * v0, ni, q1, min_boxes, thr,
                                    s.pts - 4 * q0, s.valid - q0, s.arg ? s.arg - 4 * q0 : nullptr,
                                    s.high, s.count, s.ws, ws_bytes, s.stream);
        if (rc) return rc;
        if (d_modes) if (int mrc = launch_tile_modes(s.ws, ni, d_modes, s.stream)) return mrc;
        if (nq) {
            if (h_pts) DYD_CUDA(cudaMemcpyAsync(h_pts + 4 * q0, s.pts, sizeof(double) * 4 * nq, cudaMemcpyDeviceToHost, s.stream));
            if (h_valid) DYD_CUDA(cudaMemcpyAsync(h_valid + q0, s.valid, nq, cudaMemcpyDeviceToHost, s.stream));
            if (h_arg) DYD_CUDA(cudaMemcpyAsync(h_arg + 4 * q0, s.arg, sizeof(int32_t) * 4 * nq, cudaMemcpyDeviceToHost, s.stream));
        }
        DYD_CUDA(cudaMemcpyAsync(h_high + i0, s.high, ni, cudaMemcpyDeviceToHost, s.stream));
        DYD_CUDA(cudaMemcpyAsync(h_count + i0, s.count, sizeof(int32_t) * ni, cudaMemcpyDeviceToHost, s.stream));
    }
    for (int k = 0; k < nslot; ++k) DYD_CUDA(cudaStreamSynchronize(guard.s[k].stream));
    if (d_modes) {
        unsigned long long m[3];
        DYD_CUDA(cudaMemcpy(m, d_modes, sizeof(m), cudaMemcpyDeviceToHost));
        for (int k = 0; k < 3; ++k) h_tile_modes[k] = (int64_t)m[k];
        DYD_CUDA(cudaFreeAsync(d_modes, guard.s[0].stream));
    }
    return 0;
}

extern "C" int dyd_bbox_iou_host(const int64_t* h_img_off, const int64_t* h_poly_off, const double* h_xy,
                                 int64_t n_img, int64_t min_boxes, double thr,
                                 double* h_pts, uint8_t* h_valid, int32_t* h_arg,
                                 uint8_t* h_high, int32_t* h_count, int64_t chunk_images) {
    return dyd_bbox_iou_host_ex(h_img_off, h_poly_off, h_xy, n_img, min_boxes, thr, h_pts, h_valid, h_arg, h_high, h_count,
                                chunk_images, nullptr);
}

extern "C" int dyd_dedup_host(const int64_t* h_off, const uint8_t* h_bytes, const uint8_t* h_null, int64_t n,
                              int keep_mode, uint8_t* h_keep, int64_t* h_rep) {
    DYD_REQUIRE(n >= 0, DYD_E_ARG, "negative count");
    if (n == 0) return 0;
    DYD_REQUIRE(h_off && h_keep && h_rep, DYD_E_ARG, "null pointer");
    const int64_t nbytes = h_off[n] - h_off[0];
    DYD_REQUIRE(nbytes >= 0 && (nbytes == 0 || h_bytes), DYD_E_ARG, "bad string buffer");
    cudaMemPool_t pool;
    if (int prc = host_pool(&pool)) return prc;
    cudaStream_t st;
    DYD_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    int64_t* d_off = nullptr; uint8_t* d_bytes = nullptr; uint8_t* d_null = nullptr; uint64_t* d_hash = nullptr;
    uint8_t* d_keep = nullptr; int64_t* d_rep = nullptr; void* d_ws = nullptr;
    const size_t ws_bytes = dyd_dedup_workspace_bytes(n);
    int rc = 0;
    auto body = [&]() -> int {
        DYD_CUDA(cudaMallocFromPoolAsync((void**)&d_off, sizeof(int64_t) * (n + 1), pool, st));
        DYD_CUDA(cudaMallocFromPoolAsync((void**)&d_bytes, std::max<int64_t>(nbytes, 1) + 8, pool, st));
        if (h_null) DYD_CUDA(cudaMallocFromPoolAsync((void**)&d_null, n, pool, st));
        DYD_CUDA(cudaMallocFromPoolAsync((void**)&d_hash, sizeof(uint64_t) * n, pool, st));
        DYD_CUDA(cudaMallocFromPoolAsync((void**)&d_keep, n, pool, st));
        DYD_CUDA(cudaMallocFromPoolAsync((void**)&d_rep, sizeof(int64_t) * n, pool, st));
        DYD_CUDA(cudaMallocFromPoolAsync(&d_ws, ws_bytes, pool, st));
        DYD_CUDA(cudaMemcpyAsync(d_off, h_off, sizeof(int64_t) * (n + 1), cudaMemcpyHostToDevice, st));
        if (nbytes) DYD_CUDA(cudaMemcpyAsync(d_bytes, h_bytes + h_off[0], nbytes, cudaMemcpyHostToDevice, st));
        if (h_null) DYD_CUDA(cudaMemcpyAsync(d_null, h_null, n, cudaMemcpyHostToDevice, st));
        // offsets are relative to h_bytes; the device copy starts at h_off[0]
        if (int r = dyd_hash_strings(d_off, d_bytes - h_off[0], n, d_hash, st)) return r;
        if (int r = dyd_dedup(d_hash, d_null, n, keep_mode, d_keep, d_rep, d_ws, ws_bytes, st)) return r;
        DYD_CUDA(cudaMemcpyAsync(h_keep, d_keep, n, cudaMemcpyDeviceToHost, st));
        DYD_CUDA(cudaMemcpyAsync(h_rep, d_rep, sizeof(int64_t) * n, cudaMemcpyDeviceToHost, st));
        DYD_CUDA(cudaStreamSynchronize(st));
        return 0;
    };
    rc = body();
    void* bufs[] = {d_off, d_bytes, d_null, d_hash, d_keep, d_rep, d_ws};
    for (void* b : bufs) if (b) cudaFreeAsync(b, st);
    cudaStreamSynchronize(st);
    cudaStreamDestroy(st);
    return rc;
}

extern "C" int dyd_antijoin_host(const int64_t* h_off, const uint8_t* h_bytes, const uint8_t* h_null, int64_t n,
                                 const int64_t* h_ref_off, const uint8_t* h_ref_bytes, const uint8_t* h_ref_null, int64_t n_ref,
                                 uint8_t* h_keep, int64_t* h_ref_row) {
    DYD_REQUIRE(n >= 0 && n_ref >= 0, DYD_E_ARG, "negative count");
    if (n == 0) return 0;
    DYD_REQUIRE(h_off && h_keep && h_ref_row && (n_ref == 0 || h_ref_off), DYD_E_ARG, "null pointer");
    const int64_t nbytes = h_off[n] - h_off[0];
    const int64_t rbytes = n_ref ? h_ref_off[n_ref] - h_ref_off[0] : 0;
    DYD_REQUIRE(nbytes >= 0 && rbytes >= 0 && (nbytes == 0 || h_bytes) && (rbytes == 0 || h_ref_bytes), DYD_E_ARG, "bad string buffer");
    cudaMemPool_t pool;
    if (int prc = host_pool(&pool)) return prc;
    cudaStream_t st;
    DYD_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    int64_t *d_off = nullptr, *d_roff = nullptr, *d_row = nullptr;
    uint8_t *d_bytes = nullptr, *d_rbytes = nullptr, *d_null = nullptr, *d_rnull = nullptr, *d_keep = nullptr;
    uint64_t *d_hash = nullptr, *d_rhash = nullptr;
    void* d_ws = nullptr;
    const size_t ws_bytes = dyd_antijoin_workspace_bytes(n_ref);
    auto body = [&]() -> int {
        DYD_CUDA(cudaMallocFromPoolAsync((void**)&d_off, sizeof(int64_t) * (n + 1), pool, st));
        DYD_CUDA(cudaMallocFromPoolAsync((void**)&d_bytes, std::max<int64_t>(nbytes, 1) + 8, pool, st));
        DYD_CUDA(cudaMallocFromPoolAsync((void**)&d_roff, sizeof(int64_t) * (n_ref + 1), pool, st));
        DYD_CUDA(cudaMallocFromPoolAsync((void**)&d_rbytes, std::max<int64_t>(rbytes, 1) + 8, pool, st));
        if (h_null) DYD_CUDA(cudaMallocFromPoolAsync((void**)&d_null, n, pool, st));
        if (h_ref_null && n_ref) DYD_CUDA(cudaMallocFromPoolAsync((void**)&d_rnull, n_ref, pool, st));
        DYD_CUDA(cudaMallocFromPoolAsync((void**)&d_hash, sizeof(uint64_t) * n, pool, st));
        DYD_CUDA(cudaMallocFromPoolAsync((void**)&d_rhash, sizeof(uint64_t) * std::max<int64_t>(n_ref, 1), pool, st));
        DYD_CUDA(cudaMallocFromPoolAsync((void**)&d_keep, n, pool, st));
        DYD_CUDA(cudaMallocFromPoolAsync((void**)&d_row, sizeof(int64_t) * n, pool, st));
        DYD_CUDA(cudaMallocFromPoolAsync(&d_ws, ws_bytes, pool, st));
        DYD_CUDA(cudaMemcpyAsync(d_off, h_off, sizeof(int64_t) * (n + 1), cudaMemcpyHostToDevice, st));
        if (nbytes) DYD_CUDA(cudaMemcpyAsync(d_bytes, h_bytes + h_off[0], nbytes, cudaMemcpyHostToDevice, st));
        if (h_null) DYD_CUDA(cudaMemcpyAsync(d_null, h_null, n, cudaMemcpyHostToDevice, st));
        if (n_ref) {
            DYD_CUDA(cudaMemcpyAsync(d_roff, h_ref_off, sizeof(int64_t) * (n_ref + 1), cudaMemcpyHostToDevice, st));
            if (rbytes) DYD_CUDA(cudaMemcpyAsync(d_rbytes, h_ref_bytes + h_ref_off[0], rbytes, cudaMemcpyHostToDevice, st));
            if (d_rnull) DYD_CUDA(cudaMemcpyAsync(d_rnull, h_ref_null, n_ref, cudaMemcpyHostToDevice, st));
            if (int r = dyd_hash_strings(d_roff, d_rbytes - h_ref_off[0], n_ref, d_rhash, st)) return r;
        }
        if (int r = dyd_hash_strings(d_off, d_bytes - h_off[0], n, d_hash, st)) return r;
        if (int r = dyd_antijoin(d_hash, d_null, n, d_rhash, d_rnull, n_ref, d_keep, d_row, d_ws, ws_bytes, st)) return r;
        DYD_CUDA(cudaMemcpyAsync(h_keep, d_keep, n, cudaMemcpyDeviceToHost, st));
        DYD_CUDA(cudaMemcpyAsync(h_ref_row, d_row, sizeof(int64_t) * n, cudaMemcpyDeviceToHost, st));
        DYD_CUDA(cudaStreamSynchronize(st));
        return 0;
    };
    const int rc = body();
    void* bufs[] = {d_off, d_bytes, d_roff, d_rbytes, d_null, d_rnull, d_hash, d_rhash, d_keep, d_row, d_ws};
    for (void* b : bufs) if (b) cudaFreeAsync(b, st);
    cudaStreamSynchronize(st);
    cudaStreamDestroy(st);
    return rc;
}

extern "C" int dyd_host_release(void) {
    int dev = 0;
    DYD_CUDA(cudaGetDevice(&dev));
    DYD_REQUIRE(dev >= 0 && dev < MAX_DEVICES, DYD_E_ARG, "device index out of range");
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    if (g_pools[dev]) DYD_CUDA(cudaMemPoolTrimTo(g_pools[dev], 0));
    return 0;
}

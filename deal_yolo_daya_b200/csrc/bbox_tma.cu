// Fused K1+K2, TMA-staged and warp-specialised (the sm_100a production path).
//
// One persistent CTA per SM.  The image range is cut into tiles of TILE_IMAGES images whose
// vertices are one contiguous byte range of `xy`.  Per CTA:
//
//   producer (1 thread)   waits for a free stage, then issues cp.async.bulk (TMA, UBLKCP) copies of
//                         the tile's vertex range, its poly_off slice and its img_off slice into
//                         shared memory; completion is counted in bytes on the stage's mbarrier.
//   K1 warps (12)         wait on `full`, fold polygons 8 at a time (4 lanes per polygon, 8
//                         register slots per lane, conflict-free 64-byte LDS runs), write the two
//                         corner points to HBM once and the normalised box to the stage's box
//                         array, then arrive on `ready`.
//   K2 warps (4)          wait on `ready`, run the any-pair IoU test per image straight from
//                         shared memory, write count/high, then arrive on `empty`.
//
// The compute warps never wait on HBM: the dependent img_off -> poly_off -> xy address chain is
// resolved by a descriptor pre-pass kernel and by the copy engine, and ~3 stages (~80 KB) of
// vertex data are in flight per SM.  Work inside a stage is handed out with shared-memory
// counters, so ragged polygon / image sizes do not idle warps.  Tiles that do not fit a stage
// (very large polygons) or would make a bulk copy run past the end of an offsets array take the
// fallback lane: K1 reads them with direct loads and their images are finished by the
// block-per-image kernel.
#include "kernels.cuh"

namespace dyd {

constexpr int T = TILE_IMAGES;
constexpr int STAGES = 4;
constexpr int CAP_V = 2560;                    // vertices per stage  (40 KB)
constexpr int CAP_P = 192;                     // objects per stage
constexpr int NK1 = 12, NK2 = 4;
constexpr int TMA_THREADS = 32 * (1 + NK1 + NK2);
constexpr int KG = 4, KS = 8;                  // lanes per polygon, register slots per lane
constexpr int CHUNK = 32 / KG;                 // polygons per K1 step
enum { MODE_FAST = 0, MODE_FALLBACK = 1 };

struct __align__(16) Stage {
    double2 vert[CAP_V];
    double box[CAP_P * 4];                     // (x1, y1, x2, y2) after extract_boxes' min/max
    long long poly[CAP_P + 2];                 // poly_off slice starting at object (q0 & ~1)
    long long img[T + 2];                      // img_off slice
    unsigned char bvalid[CAP_P];
    long long q0, v0;
    int tile, ni, np, pshift, mode, k1_next, k2_next, pad;
};
static_assert(sizeof(double2) * CAP_V % 16 == 0 && sizeof(double) * CAP_P * 4 % 16 == 0 &&
                  sizeof(long long) * (CAP_P + 2) % 16 == 0 && sizeof(long long) * (T + 2) % 16 == 0,
              "bulk-copy destinations must stay 16-byte aligned");

struct Smem {
    Stage st[STAGES];
    unsigned long long full[STAGES], ready[STAGES], empty[STAGES];
    __align__(16) unsigned short lut[PAIR_LUT_N + 8];
};

// ---- mbarrier / bulk-copy primitives (PTX) ------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    long long t0 = 0;
    for (unsigned spins = 0;; ++spins) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
        if (ok) return;
        if (spins == 4096) t0 = clock64();
        if (spins > 4096 && (spins & 4095) == 0 && clock64() - t0 > 6000000000LL) {   // ~3 s: a protocol bug, not load
            printf("dyd: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
            __trap();
        }
    }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---- descriptor pre-pass: resolves the img_off -> poly_off chain once per tile -------------------
__global__ void __launch_bounds__(256)
tile_desc_kernel(const int64_t* __restrict__ img_off, const int64_t* __restrict__ poly_off, int64_t n_img,
                 int64_t n_tiles, TileDesc* __restrict__ desc) {
    const int64_t k = blockIdx.x * 256LL + threadIdx.x;
    if (k >= n_tiles) return;
    const int64_t i0 = k * T, i1 = min(i0 + T, n_img);
    TileDesc d;
    d.q0 = img_off[i0]; d.q1 = img_off[i1];
    d.v0 = poly_off[d.q0]; d.v1 = poly_off[d.q1];
    desc[k] = d;
}

__device__ __forceinline__ int grab(int* counter, int lane) {
    int v = 0;
    if (lane == 0) v = atomicAdd(counter, 1);
    return __shfl_sync(FULL, v, 0);
}

template <bool ARG>
__global__ void __launch_bounds__(TMA_THREADS, 1)
fused_tma_kernel(const int64_t* __restrict__ img_off, const int64_t* __restrict__ poly_off,
                 const double2* __restrict__ xy2, const TileDesc* __restrict__ desc,
                 int64_t n_img, int64_t n_poly, int64_t n_tiles, int64_t min_boxes, double thr,
                 double* __restrict__ pts, uint8_t* __restrict__ valid, int32_t* __restrict__ arg,
                 uint8_t* __restrict__ high, int32_t* __restrict__ count, void* ws) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    load_pair_lut(sm.lut, threadIdx.x, TMA_THREADS);
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.ready[s], NK1); mbar_init(&sm.empty[s], NK2); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == 0) {
        // ================================ producer ================================
        if (lane != 0) return;
        int64_t k = blockIdx.x;
        TileDesc nxt{0, 0, 0, 0};
        if (k < n_tiles) nxt = desc[k];
        for (unsigned it = 0; k < n_tiles; k += gridDim.x, ++it) {
            const int s = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1;
            const TileDesc d = nxt;
            if (k + gridDim.x < n_tiles) nxt = desc[k + gridDim.x];          // overlaps the wait below
            mbar_wait(&sm.empty[s], ph ^ 1);
            Stage& st = sm.st[s];
            const int64_t i0 = k * T;
            const int ni = (int)min((int64_t)T, n_img - i0);
            const int64_t nv = d.v1 - d.v0, np = d.q1 - d.q0;
            const int pshift = (int)(d.q0 & 1);
            const int64_t ne = (pshift + np + 2) & ~1LL;                     // poly_off entries copied (even)
            const int64_t nie = (ni + 2) & ~1;                               // img_off entries copied (even)
            const bool fits = nv <= CAP_V && pshift + np + 1 <= CAP_P;
            const bool tail_ok = (d.q0 - pshift) + ne <= n_poly + 1 && i0 + nie <= n_img + 1;
            st.q0 = d.q0; st.v0 = d.v0; st.tile = (int)k; st.ni = ni; st.pshift = pshift;
            st.k1_next = 0; st.k2_next = 0;
            if (fits && tail_ok) {
                st.np = (int)np; st.mode = MODE_FAST;
                mbar_arrive_expect_tx(&sm.full[s], (uint32_t)(16 * nv + 8 * ne + 8 * nie));
                if (nv > 0) bulk_g2s(st.vert, xy2 + d.v0, (uint32_t)(16 * nv), &sm.full[s]);
                bulk_g2s(st.poly, poly_off + (d.q0 - pshift), (uint32_t)(8 * ne), &sm.full[s]);
                bulk_g2s(st.img, img_off + i0, (uint32_t)(8 * nie), &sm.full[s]);
            } else {
                st.np = (int)min(np, (int64_t)0x7fffffff); st.mode = MODE_FALLBACK;
                mbar_arrive(&sm.full[s]);
            }
        }
    } else if (warp <= NK1) {
        // ================================ K1: polygon -> corner points ================================
        const int gl = lane & (KG - 1), g = lane / KG;
        const unsigned gmask = ((1u << KG) - 1u) << (g * KG);
        unsigned it = 0;
        for (int64_t k = blockIdx.x; k < n_tiles; k += gridDim.x, ++it) {
            const int s = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1;
            mbar_wait(&sm.full[s], ph);
            Stage& st = sm.st[s];
            const int np = st.np, pshift = st.pshift, mode = st.mode;
            const int64_t q0 = st.q0, v0 = st.v0;
            for (;;) {
                const int c0 = grab(&st.k1_next, lane) * CHUNK;
                if (c0 >= np) break;
                const int pl = c0 + g;
                if (pl < np) {
                    const int64_t p = q0 + pl;
                    Corner c{0.0, 0.0, 0.0, 0.0};
                    CornerIdx ci{-1, -1, -1, -1};
                    int V;
                    if (mode == MODE_FAST) {
                        const long long a = st.poly[pshift + pl], b = st.poly[pshift + pl + 1];
                        V = (int)(b - a);
                        const double2* base = st.vert + (a - v0);
                        auto load = [&](int kk) { return base[kk]; };
                        if (V > 0) c = group_bbox<KG, KS, ARG>(load, V, gl, gmask, ci);
                    } else {
                        const int64_t a = __ldg(poly_off + p), b = __ldg(poly_off + p + 1);
                        const int64_t Vl = b - a;
                        V = Vl > 0x7fffffff ? 0x7fffffff : (int)Vl;
                        auto load = [&](int kk) { return ldg_stream_f64x2(xy2 + a + kk); };
                        if (V > 0) c = group_bbox<KG, KS, ARG>(load, V, gl, gmask, ci);
                    }
                    if (gl == 0) {
                        double2* o = reinterpret_cast<double2*>(pts + 4 * p);
                        stg_stream_f64x2(o, make_double2(c.mnx, c.mny));
                        stg_stream_f64x2(o + 1, make_double2(c.mxx, c.mxy));
                        valid[p] = V > 0 ? 1 : 0;
                        if (ARG) *reinterpret_cast<int4*>(arg + 4 * p) = make_int4(ci.mnx, ci.mny, ci.mxx, ci.mxy);
                        if (mode == MODE_FAST) {
                            const Box bx = box_from_points(c.mnx, c.mny, c.mxx, c.mxy);
                            double2* sb = reinterpret_cast<double2*>(st.box + 4 * pl);
                            sb[0] = make_double2(bx.x1, bx.y1); sb[1] = make_double2(bx.x2, bx.y2);
                            st.bvalid[pl] = V > 0 ? 1 : 0;
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.ready[s]);
        }
    } else {
        // ================================ K2: box count + any-pair IoU ================================
        const bool zero_hits = 0.0 >= thr;
        unsigned it = 0;
        for (int64_t k = blockIdx.x; k < n_tiles; k += gridDim.x, ++it) {
            const int s = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1;
            mbar_wait(&sm.ready[s], ph);
            Stage& st = sm.st[s];
            const int ni = st.ni, mode = st.mode;
            const int64_t q0 = st.q0, i0 = (int64_t)st.tile * T;
            for (;;) {
                const int j = grab(&st.k2_next, lane);
                if (j >= ni) break;
                const int64_t img = i0 + j;
                int n = WARP_BOX_CAP + 1, lq = 0;
                if (mode == MODE_FAST) { lq = (int)(st.img[j] - q0); n = (int)(st.img[j + 1] - st.img[j]); }
                if (n > WARP_BOX_CAP) {                   // crowded image or fallback tile
                    if (lane == 0) {
                        unsigned long long slot = atomicAdd(&reinterpret_cast<CrowdList*>(ws)->count, 1ULL);
                        crowd_ids(ws)[slot] = (int)img;
                    }
                    continue;
                }
                int n_eff = n;
                for (int base = 0; base < n; base += 32) {
                    const int jj = base + lane;
                    const unsigned m = __ballot_sync(FULL, jj < n && st.bvalid[lq + jj] == 0);
                    if (m) { n_eff = base + (__ffs(m) - 1); break; }
                }
                bool hit = false;
                if (n_eff >= min_boxes && n_eff >= 2) hit = warp_any_pair(st.box + 4 * lq, n_eff, thr, zero_hits, sm.lut, lane);
                if (lane == 0) { count[img] = n_eff; high[img] = hit ? 1 : 0; }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.empty[s]);
        }
    }
}

int launch_fused_tma(const int64_t* d_img_off, const int64_t* d_poly_off, const double* d_xy,
                     int64_t n_img, int64_t n_poly, int64_t min_boxes, double thr,
                     double* d_pts, uint8_t* d_valid, int32_t* d_arg, uint8_t* d_high, int32_t* d_count,
                     void* ws, cudaStream_t s) {
    const int64_t n_tiles = n_tiles_of(n_img);
    TileDesc* desc = tile_descs(ws, n_img);
    tile_desc_kernel<<<(unsigned)((n_tiles + 255) / 256), 256, 0, s>>>(d_img_off, d_poly_off, n_img, n_tiles, desc);
    if (int rc = launch_check("tile_desc_kernel")) return rc;
    const size_t smem = sizeof(Smem);
    const unsigned grid = (unsigned)(n_tiles < NUM_SMS ? n_tiles : NUM_SMS);
    const double2* xy2 = reinterpret_cast<const double2*>(d_xy);
    if (d_arg) {
        DYD_CUDA(cudaFuncSetAttribute(fused_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        fused_tma_kernel<true><<<grid, TMA_THREADS, smem, s>>>(d_img_off, d_poly_off, xy2, desc, n_img, n_poly, n_tiles, min_boxes, thr,
                                                               d_pts, d_valid, d_arg, d_high, d_count, ws);
    } else {
        DYD_CUDA(cudaFuncSetAttribute(fused_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        fused_tma_kernel<false><<<grid, TMA_THREADS, smem, s>>>(d_img_off, d_poly_off, xy2, desc, n_img, n_poly, n_tiles, min_boxes, thr,
                                                                d_pts, d_valid, nullptr, d_high, d_count, ws);
    }
    return launch_check("fused_tma_kernel");
}

}  // namespace dyd

// Fused K1+K2, TMA-staged (the sm_100a production path).
//
// One persistent CTA per SM, NW independent warps per CTA.  A pre-pass kernel packs consecutive
// images greedily into tiles of at most 32 objects / 6 images / TILE_CAP_V vertices (so every object
// of a tile gets its own K1 lane and every tile fits a stage) and resolves the dependent
// img_off -> poly_off -> xy address chain into 32-byte descriptors.  A tile's vertices are one
// contiguous byte range of `xy`.  Each warp owns a private shared-memory stage and an mbarrier and
// walks the tiles of its segments:
//
//   1. lane 0 issues cp.async.bulk (TMA, SASS UBLKCP) copies of the tile's vertex range, poly_off
//      slice and img_off slice into the stage; completion is counted in bytes on the mbarrier;
//   2. K1: one lane per polygon folds its vertices straight out of shared memory with the
//      reference's own left fold (strict comparisons from the first vertex on, software-pipelined
//      over two register buffers), which is CPython's min()/max() bit for bit -- ties, signed
//      zeros and NaN order need no special cases -- and writes the two corner points to HBM once
//      and the normalised box to the stage;
//   3. the next tile's copies are issued (the vertex buffer is free again), so they land during
//   4. K2 (k2_tile.cuh): per-image box counts in parallel lanes; then lane b owns box b and meets the
//      boxes of its image at circular distances 1 .. n/2 (every pair once) with an overlap pre-test;
//      survivors are spread over a small queue and evaluated densely with the full IoU arithmetic.
//
// Warps never synchronise with each other; while one waits for its copy the others compute, so the
// copy engine keeps tens of KB per SM in flight without any register cost.  A single image that
// exceeds the stage is folded with direct loads (vertices) or handed to the block-per-image kernel
// (more than 32 objects); tiles whose offset slices would make a bulk copy run past the end of an
// array read them with plain loads.
#include <cstddef>

#include "kernels.cuh"
#include "k2_tile.cuh"

namespace dyd {

#ifndef DYD_NW
#define DYD_NW 20
#endif
constexpr int NW = DYD_NW;                     // warps per CTA (1 CTA per SM)
constexpr int CAP_V = TILE_CAP_V;
constexpr int CAP_P = TILE_LANES;              // objects per tile = K1 lanes
constexpr int TM = TILE_MAX_IMAGES;
constexpr int TMA_THREADS = 32 * NW;
constexpr int IMG_SLOTS = (TM + 3) & ~1;        // img_off slice: <= TM+1 entries + alignment shift, even count
enum { MODE_FAST = 0, MODE_DIRECT = 1, MODE_DEFER = 2, MODE_END = 3 };
static_assert(TM + 2 <= IMG_SLOTS && TM <= 8 && CAP_P == 32 && SEG_IMAGES <= 255, "slice sizes");

struct __align__(16) TileInfo {                // double-buffered: the next tile is described while K2 still works
    long long img[IMG_SLOTS];                  // img_off slice starting at image (i0 & ~1)
    long long q0, v0, i0;
    int ni, np, pshift, ishift, mode, pad;
};
struct __align__(16) Stage {
    double2 vert[CAP_V];
    K2Tile k2;                                 // boxes of the tile + K2 scratch
    long long poly[CAP_P + 4];                 // poly_off slice starting at object (q0 & ~1)
    TileInfo info[2];
    unsigned char bvalid[CAP_P];
    unsigned long long bar;
    unsigned long long pad;
};
static_assert(sizeof(double2) * CAP_V % 16 == 0 && sizeof(long long) * (CAP_P + 4) % 16 == 0 && sizeof(TileInfo) % 16 == 0,
              "bulk-copy destinations must stay 16-byte aligned");

struct Smem {
    Stage st[NW];
};
static_assert(sizeof(Smem) <= 227 * 1024, "stage set exceeds the 227 KB shared memory of an SM");

// ---- mbarrier / bulk-copy primitives (PTX) ------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t addr, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(addr), "r"(parity), "r"(0x4000u) : "memory");     // suspend-time hint (ns)
    return ok != 0;
}
__device__ __noinline__ void mbar_wait_slow(uint32_t addr, uint32_t parity) {
#ifdef DYD_DEBUG_WATCHDOG                      // debug builds only: under time-slicing / MPS / a debugger a legitimate
    const long long t0 = clock64();            // wait can be arbitrarily long, and a trap poisons the caller's context
    while (!mbar_try_wait(addr, parity)) {
        if (clock64() - t0 > 6000000000LL) {
            printf("dyd: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
            __trap();
        }
    }
#else
    while (!mbar_try_wait(addr, parity)) {}
#endif
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    for (int spins = 0; spins < 64; ++spins)
        if (mbar_try_wait(addr, parity)) return;
    mbar_wait_slow(addr, parity);
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---- descriptor pre-pass: one warp per segment packs its images greedily into tiles ------------------
// Lane l fetches img_off[i_begin + l] and the poly_off entry it points to (two independent load levels
// instead of a 32-step pointer chase).  Every lane then finds, from the two prefix arrays, where the
// maximal tile starting at ITS image would end; lane 0 follows those jump pointers from image 0 (a
// handful of shared-memory reads) and lane t writes the descriptor of tile t.
static_assert(SEG_IMAGES == 32, "one lane per image of a segment");
constexpr int DESC_WARPS = 8;
__global__ void __launch_bounds__(32 * DESC_WARPS)
tile_desc_kernel(const int64_t* __restrict__ img_off, const int64_t* __restrict__ poly_off, int64_t n_img,
                 int64_t n_poly, int64_t n_seg, TileDesc* __restrict__ desc) {
    __shared__ long long sq[DESC_WARPS][SEG_IMAGES + 1], sv[DESC_WARPS][SEG_IMAGES + 1];
    __shared__ unsigned char snext[DESC_WARPS][SEG_IMAGES], sstart[DESC_WARPS][SEG_IMAGES + 1];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t s = blockIdx.x * (int64_t)DESC_WARPS + w;
    if (s >= n_seg) return;
    const int64_t i_begin = s * SEG_IMAGES;
    const int n_here = (int)min((int64_t)SEG_IMAGES, n_img - i_begin);
    {
        const int e = min(lane, n_here);
        const long long q = img_off[i_begin + e];
        sq[w][e] = q; sv[w][e] = poly_off[q];
        if (lane == 0 && n_here == SEG_IMAGES) {
            const long long qe = img_off[i_begin + SEG_IMAGES];
            sq[w][SEG_IMAGES] = qe; sv[w][SEG_IMAGES] = poly_off[qe];
        }
    }
    __syncwarp();
    if (lane < n_here) {                                       // end of the maximal tile that starts at image `lane`
        const long long q0 = sq[w][lane], v0 = sv[w][lane];
        int k = 1;
        if (sq[w][lane + 1] - q0 <= TILE_LANES && sv[w][lane + 1] - v0 <= TILE_CAP_V)
            while (k < TILE_MAX_IMAGES && lane + k < n_here && sq[w][lane + k + 1] - q0 <= TILE_LANES && sv[w][lane + k + 1] - v0 <= TILE_CAP_V) ++k;
        snext[w][lane] = (unsigned char)(lane + k);
    }
    __syncwarp();
    int cnt = 0;
    if (lane == 0) {
        for (int pos = 0; pos < n_here; pos = snext[w][pos]) sstart[w][cnt++] = (unsigned char)pos;
        sstart[w][cnt] = (unsigned char)n_here;
    }
    cnt = __shfl_sync(FULL, cnt, 0);
    __syncwarp();
    if (lane < cnt) {
        const int a = sstart[w][lane], b = sstart[w][lane + 1];
        const int64_t t_i0 = i_begin + a;
        const long long t_q0 = sq[w][a], t_v0 = sv[w][a], t_np = sq[w][b] - t_q0, t_nv = sv[w][b] - t_v0;
        const int t_ni = b - a;
        TileDesc d;
        d.q0 = t_q0; d.v0 = t_v0; d.i0 = (int)t_i0;
        d.nv = (int)min(t_nv, (long long)0x7fffffff); d.np = (short)min(t_np, (long long)32767); d.ni = (unsigned char)t_ni;
        int mode = t_np > TILE_LANES ? MODE_DEFER : (t_nv > TILE_CAP_V ? MODE_DIRECT : MODE_FAST);   // oversize = single image
        if (mode == MODE_FAST) {                               // bulk copies read whole 16-byte units: stay inside the arrays
            const int pshift = (int)(t_q0 & 1), ishift = (int)(t_i0 & 1);
            const int64_t ne = (pshift + t_np + 2) & ~1LL, nie = (ishift + t_ni + 2) & ~1LL;
            if ((t_q0 - pshift) + ne > n_poly + 1 || (t_i0 - ishift) + nie > n_img + 1) mode = MODE_DIRECT;
        }
        d.mode = (unsigned char)mode; d.cnt = lane == 0 ? (unsigned char)cnt : 0; d.pad[0] = d.pad[1] = d.pad[2] = 0;
        desc[s * SEG_IMAGES + lane] = d;
    }
}

// The reference's fold for one polygon (processor.py:256-259): running values start at vertex 0 and
// are replaced only on a strict comparison, in vertex order.  LOAD(k) returns vertex k.
template <bool ARG, typename LoadFn>
__device__ __forceinline__ Corner fold_sequential(LoadFn load, int V, CornerIdx& ci) {
    const double2 f = load(0);
    Corner c{f.x, f.y, f.x, f.y};
    ci = CornerIdx{0, 0, 0, 0};
#pragma unroll 4
    for (int k = 1; k < V; ++k) {
        const double2 v = load(k);
        if (ARG) {
            if (v.x < c.mnx) { c.mnx = v.x; ci.mnx = k; }
            if (v.x > c.mxx) { c.mxx = v.x; ci.mxx = k; }
            if (v.y < c.mny) { c.mny = v.y; ci.mny = k; }
            if (v.y > c.mxy) { c.mxy = v.y; ci.mxy = k; }
        } else {
            c.mnx = v.x < c.mnx ? v.x : c.mnx;
            c.mxx = v.x > c.mxx ? v.x : c.mxx;
            c.mny = v.y < c.mny ? v.y : c.mny;
            c.mxy = v.y > c.mxy ? v.y : c.mxy;
        }
    }
    return c;
}

// The same fold over a polygon staged in shared memory, software-pipelined: the loads of the next
// DYD_K1_PIPE vertices are in flight while the current ones are compared (a shared-memory round trip is
// ~30 cycles, one group of compares ~50).  `base` may be read up to 3 * DYD_K1_PIPE entries past the
// polygon's end: that is still inside the stage (the caller guarantees it) and the values are not used.
#ifndef DYD_K1_PIPE
#define DYD_K1_PIPE 2
#endif
template <bool ARG>
__device__ __forceinline__ void fold_step(Corner& c, CornerIdx& ci, const double2 v, int k) {
    if (ARG) {
        if (v.x < c.mnx) { c.mnx = v.x; ci.mnx = k; }
        if (v.x > c.mxx) { c.mxx = v.x; ci.mxx = k; }
        if (v.y < c.mny) { c.mny = v.y; ci.mny = k; }
        if (v.y > c.mxy) { c.mxy = v.y; ci.mxy = k; }
    } else {
        c.mnx = v.x < c.mnx ? v.x : c.mnx;
        c.mxx = v.x > c.mxx ? v.x : c.mxx;
        c.mny = v.y < c.mny ? v.y : c.mny;
        c.mxy = v.y > c.mxy ? v.y : c.mxy;
    }
}
template <bool ARG>
__device__ __forceinline__ Corner fold_staged(const double2* base, int V, CornerIdx& ci) {
    constexpr int P = DYD_K1_PIPE;
    const double2 f = base[0];
    Corner c{f.x, f.y, f.x, f.y};
    ci = CornerIdx{0, 0, 0, 0};
    double2 a[P], b[P];                            // two register buffers used alternately: no copies
#pragma unroll
    for (int u = 0; u < P; ++u) a[u] = base[1 + u];
    int k = 1;
    for (; k + 2 * P <= V; k += 2 * P) {
#pragma unroll
        for (int u = 0; u < P; ++u) b[u] = base[k + P + u];
#pragma unroll
        for (int u = 0; u < P; ++u) fold_step<ARG>(c, ci, a[u], k + u);
#pragma unroll
        for (int u = 0; u < P; ++u) a[u] = base[k + 2 * P + u];
#pragma unroll
        for (int u = 0; u < P; ++u) fold_step<ARG>(c, ci, b[u], k + P + u);
    }
    // fewer than 2 P vertices left; a[] holds base[k .. k+P-1]
    if (k + P < V) {
#pragma unroll
        for (int u = 0; u < P; ++u) b[u] = base[k + P + u];
    }
#pragma unroll
    for (int u = 0; u < P; ++u) if (k + u < V) fold_step<ARG>(c, ci, a[u], k + u);
#pragma unroll
    for (int u = 0; u < P - 1; ++u) if (k + P + u < V) fold_step<ARG>(c, ci, b[u], k + P + u);
    return c;
}

// Diagnostics: globaltimer at which each CTA of the last fused launch started / finished (2 stores per CTA).
__device__ unsigned long long g_cta_times[2 * NUM_SMS];
__device__ __forceinline__ unsigned long long globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

template <bool ARG>
__global__ void __launch_bounds__(TMA_THREADS, 1)
fused_tma_kernel(const int64_t* __restrict__ img_off, const int64_t* __restrict__ poly_off,
                 const double2* __restrict__ xy2, const TileDesc* __restrict__ desc,
                 int64_t n_img, int64_t n_poly, int64_t n_seg, int64_t min_boxes, double thr,
                 double* __restrict__ pts, uint8_t* __restrict__ valid, int32_t* __restrict__ arg,
                 uint8_t* __restrict__ high, int32_t* __restrict__ count, void* ws) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    Stage& st = sm.st[warp];

    if (lane == 0) mbar_init(&st.bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (threadIdx.x == 0 && blockIdx.x < NUM_SMS) g_cta_times[2 * blockIdx.x] = globaltimer();
    __syncthreads();

    const bool zero_hits = 0.0 >= thr;
    uint32_t phase = 0;
    // Segments are claimed dynamically: every warp starts on segment (block, warp) and takes further ones from
    // an atomic counter (CrowdList::next_seg, zeroed by the entry point).  A CTA that becomes resident late --
    // another stream's kernels held its SM -- or that runs slower simply claims fewer segments, so the kernel
    // tolerates concurrent work on the GPU and skewed tables.  The claim for the segment after next is issued
    // one segment ahead; its latency is hidden behind a whole segment of tiles.
    unsigned long long* const seg_counter = &reinterpret_cast<CrowdList*>(ws)->next_seg;
    const int64_t dyn_base = (int64_t)gridDim.x * NW;
    int64_t claimed = 0;

    // Lane 0 walks the descriptors of this warp's segments two tiles ahead of the tile being
    // processed: the load issued while one tile is filled is consumed a full tile later.
    // The descriptors stay in registers as the two raw 16-byte words they were loaded as and are taken
    // apart only when consumed: unpacking at load time would make the warp wait for the load it just issued.
    struct Cursor { int64_t seg; int j, cnt; };
    struct Raw { ulonglong2 a, b; };               // a = (q0, v0), b = (i0 | nv << 32, np | ni << 16 | mode << 24 | cnt << 32)
    static_assert(sizeof(TileDesc) == 32 && offsetof(TileDesc, i0) == 16 && offsetof(TileDesc, nv) == 20 && offsetof(TileDesc, np) == 24 &&
                  offsetof(TileDesc, ni) == 26 && offsetof(TileDesc, mode) == 27 && offsetof(TileDesc, cnt) == 28, "TileDesc layout");
    auto load_raw = [&](int64_t idx) {
        const ulonglong2* q = reinterpret_cast<const ulonglong2*>(desc + idx);
        return Raw{__ldg(q), __ldg(q + 1)};
    };
    auto cnt_of = [](const Raw& r) { return (int)((r.b.y >> 32) & 0xff); };
    Cursor c_nxt{(int64_t)blockIdx.x * NW + warp, 0, 0}, c_far{0, 0, 0};
    Raw nxt{}, far{};
    bool has_nxt = false, has_far = false;
    auto advance = [&](const Cursor& c) {
        if (c.j + 1 < c.cnt) return Cursor{c.seg, c.j + 1, c.cnt};
        const Cursor r{claimed, 0, 0};
        if (claimed < n_seg) claimed = dyn_base + (int64_t)atomicAdd(seg_counter, 1ULL);
        return r;
    };
    if (lane == 0) {
        claimed = dyn_base + (int64_t)atomicAdd(seg_counter, 1ULL);
        has_nxt = c_nxt.seg < n_seg;
        if (has_nxt) {
            nxt = load_raw(c_nxt.seg * SEG_IMAGES);
            c_nxt.cnt = cnt_of(nxt);
            c_far = advance(c_nxt);
            has_far = c_far.seg < n_seg;
            if (has_far) far = load_raw(c_far.seg * SEG_IMAGES + c_far.j);
        }
    }

    // Stage fill, executed by lane 0 only: describe the next tile in `ti` and start its copies.
    auto fill = [&](TileInfo& ti) {
        if (!has_nxt) { ti.mode = MODE_END; return; }
        const Raw r = nxt;
        nxt = far; c_nxt = c_far; has_nxt = has_far;
        if (has_nxt) {
            if (c_nxt.j == 0) c_nxt.cnt = cnt_of(nxt);
            c_far = advance(c_nxt);
            has_far = c_far.seg < n_seg;
            if (has_far) far = load_raw(c_far.seg * SEG_IMAGES + c_far.j);
        }
        const int64_t d_q0 = (int64_t)r.a.x, d_v0 = (int64_t)r.a.y;
        const int64_t i0 = (int64_t)(int)(r.b.x & 0xffffffffu);
        const int d_nv = (int)(r.b.x >> 32), d_np = (int)(short)(r.b.y & 0xffff);
        const int ni = (int)((r.b.y >> 16) & 0xff), d_mode = (int)((r.b.y >> 24) & 0xff);
        const int pshift = (int)(d_q0 & 1), ishift = (int)(i0 & 1);
        ti.q0 = d_q0; ti.v0 = d_v0; ti.i0 = i0; ti.ni = ni; ti.pshift = pshift; ti.ishift = ishift; ti.np = d_np; ti.mode = d_mode;
        if (d_mode == MODE_FAST) {
            const uint32_t ne = (uint32_t)(pshift + d_np + 2) & ~1u;         // poly_off entries copied (even count)
            const uint32_t nie = (uint32_t)(ishift + ni + 2) & ~1u;          // img_off entries copied (even count)
            mbar_arrive_expect_tx(&st.bar, 16u * (uint32_t)d_nv + 8u * ne + 8u * nie);
            if (d_nv > 0) bulk_g2s(st.vert, xy2 + d_v0, 16u * (uint32_t)d_nv, &st.bar);
            bulk_g2s(st.poly, poly_off + (d_q0 - pshift), 8u * ne, &st.bar);
            bulk_g2s(ti.img, img_off + (i0 - ishift), 8u * nie, &st.bar);
        } else {                                                              // offsets by plain loads
            for (int j = 0; j <= ni; ++j) ti.img[ishift + j] = __ldg(img_off + i0 + j);
            const long long np = ti.img[ishift + ni] - d_q0;
            ti.np = (int)min(np, (long long)0x7fffffff);
        }
    };

    if (lane == 0) fill(st.info[0]);
    for (unsigned it = 0;; ++it) {
        TileInfo& ti = st.info[it & 1];
        __syncwarp();
        const int mode = ti.mode;
        if (mode == MODE_END) break;
        const int np = ti.np, pshift = ti.pshift, ishift = ti.ishift, ni = ti.ni;
        const int64_t q0 = ti.q0, v0 = ti.v0, i0 = ti.i0;
        if (mode == MODE_FAST) {
            mbar_wait(&st.bar, phase);
            phase ^= 1;
        }
        // ---------------- K1: one lane per polygon ----------------
        bool nan_box = false;
        for (int base = 0; base < np; base += 32) {
            const int pl = base + lane;
            if (pl >= np) continue;
            const int64_t p = q0 + pl;
            Corner c{0.0, 0.0, 0.0, 0.0};
            CornerIdx ci{-1, -1, -1, -1};
            int V;
            if (mode == MODE_FAST) {
                const long long a = st.poly[pshift + pl], b = st.poly[pshift + pl + 1];
                V = (int)(b - a);
                const double2* base = st.vert + (a - v0);
                if (V > 0) c = fold_staged<ARG>(base, V, ci);
            } else {
                const int64_t a = __ldg(poly_off + p), b = __ldg(poly_off + p + 1);
                const int64_t Vl = b - a;
                V = Vl > 0x7fffffff ? 0x7fffffff : (int)Vl;
                const double2* base = xy2 + a;
                if (V > 0) c = fold_sequential<ARG>([&](int kk) { return ldg_f64x2(base + kk); }, V, ci);
            }
            double2* o = reinterpret_cast<double2*>(pts + 4 * p);
            stg_stream_f64x2(o, make_double2(c.mnx, c.mny));
            stg_stream_f64x2(o + 1, make_double2(c.mxx, c.mxy));
            valid[p] = V > 0 ? 1 : 0;
            if (ARG) *reinterpret_cast<int4*>(arg + 4 * p) = make_int4(ci.mnx, ci.mny, ci.mxx, ci.mxy);
            if (mode != MODE_DEFER) {
                nan_box |= (c.mnx != c.mnx) | (c.mny != c.mny) | (c.mxx != c.mxx) | (c.mxy != c.mxy);
                const Box bx = box_from_points(c.mnx, c.mny, c.mxx, c.mxy);
                st.k2.box_lo[pl] = make_double2(bx.x1, bx.y1); st.k2.box_hi[pl] = make_double2(bx.x2, bx.y2);
                st.bvalid[pl] = V > 0 ? 1 : 0;
            }
        }
        __syncwarp();
        // The vertex / offset buffers are free again: start the next tile's copies now so that they
        // land while K2 works on this tile's boxes.
        if (lane == 0) fill(st.info[(it + 1) & 1]);

        // ---------------- K2: box counts + any-pair IoU, flattened over the tile's images ----------------
        if (mode == MODE_DEFER) {
            if (lane < ni) {
                unsigned long long slot = atomicAdd(&reinterpret_cast<CrowdList*>(ws)->count, 1ULL);
                crowd_ids(ws)[slot] = (int)(i0 + lane);
            }
            continue;
        }
        const bool exact_pre = __any_sync(FULL, nan_box);
        const unsigned inv = __ballot_sync(FULL, lane < np && st.bvalid[lane] == 0);   // bit p: object p is a null bbox
        int my_a = 0, my_n = 0, my_ne;
        if (lane < ni) {
            my_a = (int)(ti.img[ishift + lane] - q0);
            my_n = (int)(ti.img[ishift + lane + 1] - ti.img[ishift + lane]);
        }
        const unsigned hits = k2_tile_any<TM>(st.k2, inv, exact_pre, my_a, my_n, ni, np, min_boxes, thr, zero_hits, lane, my_ne);
        if (lane < ni) { count[i0 + lane] = my_ne; high[i0 + lane] = (hits >> lane) & 1u; }
        __syncwarp();                              // st.k2 scratch is rewritten by the next tile
    }
    if (threadIdx.x == 0 && blockIdx.x < NUM_SMS) atomicMax(&g_cta_times[2 * blockIdx.x + 1], globaltimer());
}

// Diagnostics: how many tiles of each mode the pre-pass produced (the descriptors stay in the workspace).
__global__ void tile_modes_kernel(const TileDesc* __restrict__ desc, int64_t n_seg, unsigned long long* counts) {
    const int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (s >= n_seg) return;
    const int cnt = desc[s * SEG_IMAGES].cnt;
    unsigned c[3] = {0, 0, 0};
    for (int j = 0; j < cnt; ++j) { const int m = desc[s * SEG_IMAGES + j].mode; if (m < 3) ++c[m]; }
    for (int m = 0; m < 3; ++m) if (c[m]) atomicAdd(&counts[m], (unsigned long long)c[m]);
}
int fused_cta_times(unsigned long long* h_out, int n) {
    if (n > 2 * NUM_SMS) n = 2 * NUM_SMS;
    DYD_CUDA(cudaMemcpyFromSymbol(h_out, g_cta_times, sizeof(unsigned long long) * n));
    return 0;
}
int launch_tile_modes(const void* ws, int64_t n_img, unsigned long long* d_counts3, cudaStream_t s) {
    const int64_t n_seg = n_segments_of(n_img);
    if (n_seg == 0) return 0;
    const TileDesc* desc = tile_descs(const_cast<void*>(ws), n_img);
    tile_modes_kernel<<<(unsigned)((n_seg + 255) / 256), 256, 0, s>>>(desc, n_seg, d_counts3);
    return launch_check("tile_modes_kernel");
}

int launch_fused_tma(const int64_t* d_img_off, const int64_t* d_poly_off, const double* d_xy,
                     int64_t n_img, int64_t n_poly, int64_t min_boxes, double thr,
                     double* d_pts, uint8_t* d_valid, int32_t* d_arg, uint8_t* d_high, int32_t* d_count,
                     void* ws, int max_ctas, cudaEvent_t prepass_done, cudaStream_t s) {
    const int64_t n_seg = n_segments_of(n_img);
    TileDesc* desc = tile_descs(ws, n_img);
    tile_desc_kernel<<<(unsigned)((n_seg + DESC_WARPS - 1) / DESC_WARPS), 32 * DESC_WARPS, 0, s>>>(d_img_off, d_poly_off, n_img, n_poly, n_seg, desc);
    if (int rc = launch_check("tile_desc_kernel")) return rc;
    if (prepass_done) DYD_CUDA(cudaEventRecord(prepass_done, s));
    const size_t smem = sizeof(Smem);
    const int64_t want = (n_seg + NW - 1) / NW;
    const int64_t cap = max_ctas > 0 && max_ctas < NUM_SMS ? max_ctas : NUM_SMS;   // < 148: leave SMs to concurrent streams
    const unsigned grid = (unsigned)(want < cap ? want : cap);
    const double2* xy2 = reinterpret_cast<const double2*>(d_xy);
    if (d_arg) {
        DYD_CUDA(cudaFuncSetAttribute(fused_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        fused_tma_kernel<true><<<grid, TMA_THREADS, smem, s>>>(d_img_off, d_poly_off, xy2, desc, n_img, n_poly, n_seg, min_boxes, thr,
                                                               d_pts, d_valid, d_arg, d_high, d_count, ws);
    } else {
        DYD_CUDA(cudaFuncSetAttribute(fused_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        fused_tma_kernel<false><<<grid, TMA_THREADS, smem, s>>>(d_img_off, d_poly_off, xy2, desc, n_img, n_poly, n_seg, min_boxes, thr,
                                                                d_pts, d_valid, nullptr, d_high, d_count, ws);
    }
    return launch_check("fused_tma_kernel");
}

}  // namespace dyd

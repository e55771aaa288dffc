// Declarations shared by bbox_iou.cu (direct-load kernels, entry points) and bbox_tma.cu
// (TMA-staged fused kernel with warp-private stages).
#pragma once
#include "bbox_core.cuh"

namespace dyd {

// ---- workspace layout of the K2 / fused entry points -------------------------------------
//   [CrowdList header 16 B][int32 image ids, n_img][pad to 16 B][TileDesc, n_tiles]
struct CrowdList {
    unsigned long long count;      // number of deferred images
    unsigned long long next_seg;   // fused kernel: dynamic segment counter (both zeroed by the entry point)
};
__host__ __device__ __forceinline__ int* crowd_ids(void* ws) {
    return reinterpret_cast<int*>(reinterpret_cast<char*>(ws) + sizeof(CrowdList));
}
inline size_t crowd_list_bytes(int64_t n_img) {
    return (sizeof(CrowdList) + sizeof(int) * (size_t)n_img + 15) & ~(size_t)15;
}

// Tiles of the fused kernel: a pre-pass packs consecutive images greedily into tiles of at most
// TILE_LANES objects (one K1 lane each), TILE_MAX_IMAGES images and TILE_CAP_V vertices.  Packing
// restarts every SEG_IMAGES images so that segments are independent (one pre-pass warp each, one
// warp of the main kernel each); a segment's descriptors live at desc[seg * SEG_IMAGES ...].
#ifndef DYD_SEG_IMAGES
#define DYD_SEG_IMAGES 32
#endif
#ifndef DYD_TILE_MAX_IMAGES
#define DYD_TILE_MAX_IMAGES 6
#endif
#ifndef DYD_TILE_CAP_V
#define DYD_TILE_CAP_V 616
#endif
constexpr int SEG_IMAGES = DYD_SEG_IMAGES;
constexpr int TILE_LANES = 32;
constexpr int TILE_MAX_IMAGES = DYD_TILE_MAX_IMAGES;
constexpr int TILE_CAP_V = DYD_TILE_CAP_V;   // vertices staged per tile (9.6 KB)
struct TileDesc {                            // 32 bytes
    long long q0;                            // first object  img_off[i0]
    long long v0;                            // first vertex  poly_off[q0]
    int i0;                                  // first image
    int nv;                                  // vertices (clamped to INT_MAX)
    short np;                                // objects (clamped; exact for fast / direct tiles)
    unsigned char ni;                        // images
    unsigned char mode;                      // MODE_*
    unsigned char cnt;                       // tiles of the segment (valid in the segment's first descriptor)
    unsigned char pad[3];
};
static_assert(sizeof(TileDesc) == 32, "descriptor layout");
inline int64_t n_segments_of(int64_t n_img) { return (n_img + SEG_IMAGES - 1) / SEG_IMAGES; }
inline size_t tile_desc_bytes(int64_t n_img) { return sizeof(TileDesc) * (size_t)n_segments_of(n_img) * SEG_IMAGES + 16; }
inline TileDesc* tile_descs(void* ws, int64_t n_img) {
    return reinterpret_cast<TileDesc*>(reinterpret_cast<char*>(ws) + crowd_list_bytes(n_img));
}

int launch_fused_tma(const int64_t* d_img_off, const int64_t* d_poly_off, const double* d_xy,
                     int64_t n_img, int64_t n_poly, int64_t min_boxes, double thr,
                     double* d_pts, uint8_t* d_valid, int32_t* d_arg, uint8_t* d_high, int32_t* d_count,
                     void* ws, int max_ctas, cudaEvent_t prepass_done, cudaStream_t s);
int fused_cta_times(unsigned long long* h_out, int n);

// [fast, direct, defer] tile counts of the descriptors a fused call left in its workspace (diagnostics)
int launch_tile_modes(const void* ws, int64_t n_img, unsigned long long* d_counts3, cudaStream_t s);

int launch_crowd(const int64_t* d_img_off, const double* d_pts, const uint8_t* d_valid, int64_t min_boxes,
                 double thr, uint8_t* d_high, int32_t* d_count, void* ws, cudaStream_t s);

}  // namespace dyd

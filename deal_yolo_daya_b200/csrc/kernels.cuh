// Declarations shared by bbox_iou.cu (direct-load kernels, entry points) and bbox_tma.cu
// (TMA-staged warp-specialised fused kernel).
#pragma once
#include "bbox_core.cuh"

namespace dyd {

// ---- workspace layout of the K2 / fused entry points -------------------------------------
//   [CrowdList header 16 B][int32 image ids, n_img][pad to 16 B][TileDesc, n_tiles]
struct CrowdList {
    unsigned long long count;   // number of deferred images
    unsigned long long pad;
};
__host__ __device__ __forceinline__ int* crowd_ids(void* ws) {
    return reinterpret_cast<int*>(reinterpret_cast<char*>(ws) + sizeof(CrowdList));
}
inline size_t crowd_list_bytes(int64_t n_img) {
    return (sizeof(CrowdList) + sizeof(int) * (size_t)n_img + 15) & ~(size_t)15;
}

constexpr int TILE_IMAGES = 3;               // images per staged tile (~24 objects: one lane each)
struct TileDesc {                            // 32 bytes, written by the descriptor pre-pass
    long long q0;                            // first object  img_off[i0]
    long long v0;                            // first vertex  poly_off[q0]
    int np;                                  // objects of the tile (clamped to INT_MAX)
    int nv;                                  // vertices of the tile (clamped to INT_MAX)
    int mode;                                // how the fused kernel handles the tile (MODE_*)
    int pad;
};
inline int64_t n_tiles_of(int64_t n_img) { return (n_img + TILE_IMAGES - 1) / TILE_IMAGES; }
inline size_t tile_desc_bytes(int64_t n_img) { return sizeof(TileDesc) * (size_t)n_tiles_of(n_img) + 16; }
inline TileDesc* tile_descs(void* ws, int64_t n_img) {
    return reinterpret_cast<TileDesc*>(reinterpret_cast<char*>(ws) + crowd_list_bytes(n_img));
}

int launch_fused_tma(const int64_t* d_img_off, const int64_t* d_poly_off, const double* d_xy,
                     int64_t n_img, int64_t n_poly, int64_t min_boxes, double thr,
                     double* d_pts, uint8_t* d_valid, int32_t* d_arg, uint8_t* d_high, int32_t* d_count,
                     void* ws, cudaStream_t s);

int launch_crowd(const int64_t* d_img_off, const double* d_pts, const uint8_t* d_valid, int64_t min_boxes,
                 double thr, uint8_t* d_high, int32_t* d_count, void* ws, cudaStream_t s);

}  // namespace dyd

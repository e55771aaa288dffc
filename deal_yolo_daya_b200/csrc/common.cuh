// Shared device/host helpers for libdyd.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/dyd.h"

namespace dyd {

constexpr unsigned FULL = 0xffffffffu;
constexpr int NUM_SMS = 148;  // B200

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define DYD_CUDA(expr)                                       \
    do {                                                     \
        cudaError_t _e = (expr);                             \
        if (_e != cudaSuccess) return ::dyd::cuda_fail(_e, #expr); \
    } while (0)

#define DYD_REQUIRE(cond, code, msg)                         \
    do {                                                     \
        if (!(cond)) { ::dyd::set_error("%s: %s", __func__, msg); return (code); } \
    } while (0)

void count_launch();
inline int launch_check(const char* what) {           // called once after every kernel launch of the library
    count_launch();
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : cuda_fail(e, what);
}

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// ---- streaming loads / stores (read-once data: keep it out of L1) -------------
__device__ __forceinline__ double2 ldg_stream_f64x2(const double2* p) {
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ double2 ldg_f64x2(const double2* p) {
    double2 r;
    asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream_f64x2(double2* p, double2 v) {
    asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}

__device__ __forceinline__ double shfl_xor_f64(unsigned mask, double v, int lane_mask, int width) {
    return __shfl_xor_sync(mask, v, lane_mask, width);
}

// CPython's builtin min(a, b) / max(a, b): the second argument wins only on a strict comparison.
__device__ __forceinline__ double pymin(double a, double b) { return b < a ? b : a; }
__device__ __forceinline__ double pymax(double a, double b) { return b > a ? b : a; }

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

}  // namespace dyd

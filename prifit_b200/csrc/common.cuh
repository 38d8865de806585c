// Shared device/host helpers for the prifit_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/prifit_b200.h"

#define PRIFIT_LO (-13.0f)   // guard_exp clamp, reference src/guard.py:6-11
#define PRIFIT_HI (75.0f)

void prifit_set_error(const char* fmt, ...);

#define PF_CHECK_ARG(cond, code, msg)                                        \
    do { if (!(cond)) { prifit_set_error("%s: %s", __func__, msg); return (code); } } while (0)

#define PF_CUDA(call)                                                        \
    do { cudaError_t e__ = (call);                                           \
         if (e__ != cudaSuccess) {                                           \
             prifit_set_error("%s: %s -> %s", __func__, #call, cudaGetErrorString(e__)); \
             return (int)e__; } } while (0)

#define PF_LAUNCH_CHECK()                                                    \
    do { cudaError_t e__ = cudaGetLastError();                               \
         if (e__ != cudaSuccess) {                                           \
             prifit_set_error("%s: launch failed -> %s", __func__, cudaGetErrorString(e__)); \
             return (int)e__; } } while (0)

static inline cudaStream_t pf_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

#ifdef __CUDACC__
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum of NV values per thread; result valid in every thread.  `red` needs NV*32 floats.
// Deterministic (fixed shuffle tree, fixed warp order).
template <int NV>
__device__ __forceinline__ void block_sum(float (&v)[NV], float* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) red[i * 32 + warp] = v[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        float a = 0.f;
        for (int w = 0; w < nwarp; ++w) a += red[i * 32 + w];
        v[i] = a;
    }
}

// order-preserving float <-> uint mapping (for radix select and atomic max on floats)
__device__ __forceinline__ uint32_t float_to_ordered(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(uint32_t u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// exp(clamp(x, -13, 75)) -- reference src/guard.py:6-11
__device__ __forceinline__ float guard_expf(float x) { return expf(fminf(fmaxf(x, PRIFIT_LO), PRIFIT_HI)); }
#endif

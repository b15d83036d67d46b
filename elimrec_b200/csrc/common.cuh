// Shared helpers for libelimrec_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/elimrec_b200.h"

#define ELIMREC_API extern "C" __attribute__((visibility("default")))

void elimrec_set_error(const char* fmt, ...);

#define ER_CHECK_ARG(cond, msg)                                   \
    do {                                                          \
        if (!(cond)) {                                            \
            elimrec_set_error("%s: %s", __func__, msg);           \
            return -1;                                            \
        }                                                         \
    } while (0)

#define ER_LAUNCH_CHECK()                                                                   \
    do {                                                                                    \
        cudaError_t e__ = cudaGetLastError();                                               \
        if (e__ != cudaSuccess) {                                                           \
            elimrec_set_error("%s: launch failed: %s", __func__, cudaGetErrorString(e__));  \
            return -3;                                                                      \
        }                                                                                   \
    } while (0)

static inline cudaStream_t er_stream(elimrec_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float4 ld_nc_f4(const float4* p) { return __ldg(p); }
__device__ __forceinline__ float4 ld_cg_f4(const float4* p) { return __ldcg(p); }

__device__ __forceinline__ void fma4(float4& a, float w, const float4& v) {
    a.x = fmaf(w, v.x, a.x);
    a.y = fmaf(w, v.y, a.y);
    a.z = fmaf(w, v.z, a.z);
    a.w = fmaf(w, v.w, a.w);
}
__device__ __forceinline__ void add4(float4& a, const float4& v) {
    a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
}

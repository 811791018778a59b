// Shared helpers for the clipself_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "../../include/clipself_b200.h"

namespace cs {

// ---------------------------------------------------------------------------------------------
// Error plumbing: every extern "C" entry returns 0 on success, else a CS_ERR_* code and leaves a
// message retrievable through cs_last_error() (thread local).
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);

#define CS_CHECK_ARG(cond, ...)                                   \
    do {                                                          \
        if (!(cond)) {                                            \
            ::cs::set_error(__VA_ARGS__);                         \
            return CS_ERR_INVALID_ARGUMENT;                       \
        }                                                         \
    } while (0)

#define CS_CUDA(call)                                                                  \
    do {                                                                               \
        cudaError_t e_ = (call);                                                       \
        if (e_ != cudaSuccess) {                                                       \
            ::cs::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,       \
                            cudaGetErrorString(e_));                                   \
            return CS_ERR_CUDA;                                                        \
        }                                                                              \
    } while (0)

#define CS_LAUNCH_CHECK() CS_CUDA(cudaGetLastError())

static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
int num_sms();

// ---------------------------------------------------------------------------------------------
// Device helpers
// ---------------------------------------------------------------------------------------------
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

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t u) {
    __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
    return __bfloat1622float2(v);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

}  // namespace cs

// common.cuh -- shared helpers for the sm_100a kernels of libmlsp_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mlsp_b200.h"

#define MLSP_FULL 0xffffffffu

namespace mlsp {

// ---- error plumbing (thread-local message, integer code at the ABI) ----------------------------
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);

#define MLSP_REQUIRE(cond, code, ...)          \
    do {                                       \
        if (!(cond)) {                         \
            ::mlsp::set_error(__VA_ARGS__);    \
            return (code);                     \
        }                                      \
    } while (0)

#define MLSP_CUDA(call)                                              \
    do {                                                             \
        cudaError_t e__ = (call);                                    \
        if (e__ != cudaSuccess) return ::mlsp::cuda_fail(e__, #call); \
    } while (0)

#define MLSP_LAUNCH_CHECK(name)                                        \
    do {                                                               \
        cudaError_t e__ = cudaGetLastError();                          \
        if (e__ != cudaSuccess) return ::mlsp::cuda_fail(e__, name);   \
    } while (0)

static inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int sm_count();

// ---- device helpers ---------------------------------------------------------------------------
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// warp-wide float min/max in one instruction (redux.sync.*.f32 is an sm_100a addition)
__device__ __forceinline__ float warp_min_f32(float v)
{
    float r;
    asm volatile("redux.sync.min.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ float warp_max_f32(float v)
{
    float r;
    asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
    return r;
}

// order-preserving map float -> uint32 (ascending), -0 canonicalised to +0 by the caller
__device__ __forceinline__ uint32_t f32_orderable(float f)
{
    uint32_t u = __float_as_uint(f);
    return u ^ ((u >> 31) ? 0xffffffffu : 0x80000000u);
}

// streaming (evict-first) 128-bit store for write-once outputs
__device__ __forceinline__ void st_stream_f4(float4 *p, float4 v) { __stcs(p, v); }

}  // namespace mlsp

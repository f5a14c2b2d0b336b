// Shared helpers for the cone_b200 CUDA library (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/cone_b200.h"

namespace cone {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

// every kernel launch goes through this: counts the launch and surfaces launch errors
#define CONE_LAUNCH_CHECK(name)                                             \
    do {                                                                    \
        ::cone::count_launch();                                             \
        cudaError_t e__ = cudaGetLastError();                               \
        if (e__ != cudaSuccess) {                                           \
            ::cone::set_error("%s: %s", name, cudaGetErrorString(e__));     \
            return CONE_ERR_CUDA;                                           \
        }                                                                   \
    } while (0)

#define CONE_CUDA(call)                                                                  \
    do {                                                                                 \
        cudaError_t e__ = (call);                                                        \
        if (e__ != cudaSuccess) {                                                        \
            ::cone::set_error("%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return CONE_ERR_CUDA;                                                        \
        }                                                                                \
    } while (0)

#define CONE_TRY(call)            \
    do {                          \
        int r__ = (call);         \
        if (r__ != CONE_OK) return r__; \
    } while (0)

#define CONE_REQUIRE(cond, ...)            \
    do {                                   \
        if (!(cond)) {                     \
            ::cone::set_error(__VA_ARGS__); \
            return CONE_ERR_INVALID;       \
        }                                  \
    } while (0)

static inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

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

// Optional per-kernel timing (cone_profile_enable): CUDA events around each launch on the launching stream,
// aggregated per category.  Off by default: a scope is then a no-op.
enum ProfCat { P_GEMM_FP32 = 0, P_GEMM_TC, P_ENC_ATTN, P_DEC_ATTN, P_LAYERNORM, P_ROWOPS, P_SCORES, P_RANKLIST, P_POOL,
               P_NMS, P_CONVERT, P_ENC_TAIL, P_ENC_TAIL_G, P_COUNT };
struct ProfScope {
    ProfScope(cudaStream_t s, ProfCat cat, double flops = 0.0, double bytes = 0.0);
    ~ProfScope();
    int slot;
    cudaStream_t stream;
};

// 148 SMs on B200; grids of grid-stride kernels are sized in multiples of this
constexpr int kNumSMs = 148;

}  // namespace cone

// Tensor-core (tcgen05 + TMA + TMEM) GEMM path, CONE_PREC_TC.  See tc_gemm.cu.
#pragma once
#include "common.cuh"

namespace cone {

struct TcWeights;  // bf16 weight copies + TMA descriptors, owned by the cone_weights handle

int tc_weights_create(TcWeights** out, cudaStream_t s);
void tc_weights_destroy(TcWeights* t);
int tc_weights_refresh(TcWeights* t, cudaStream_t s);  // after the fp32 master weights changed in place
// per-call activation staging (bf16 copy of A) lives in the caller's workspace
size_t tc_scratch_bytes(int64_t max_rows, int max_k);
void tc_set_scratch(TcWeights* t, void* scratch, size_t bytes);
bool tc_gemm_supported(int64_t M, int N, int K);
// y[M,N] = epi(x[M,K] * W[N,K]^T + b (+R)), bf16 operands, fp32 accumulate, fp32 output
int tc_gemm(TcWeights* t, const float* x, int64_t ldx, int64_t M, const float* W, const float* b, int N, int K, float* y,
            int64_t ldy, int relu, const float* R, int64_t ldr, cudaStream_t s);

// General form: fp16 operand A already in memory; optional fp32 / fp16 outputs (written by TMA stores); fused bias,
// residual (fp16 via TMA, or fp32), ReLU and LayerNorm over the row (needs N == 256).  fp16 pointers are
// passed as uint16_t*.  The output may alias the fp16 residual (in-place residual stream).
struct TcGemmArgs {
    const uint16_t* A16 = nullptr;
    int64_t lda = 0, M = 0;
    const float* W = nullptr;     // fp32 master weight [N, K]; its fp16 copy is cached in TcWeights
    const float* bias = nullptr;
    int N = 0, K = 0;
    float* C32 = nullptr;
    int64_t ldc32 = 0;
    uint16_t* C16 = nullptr;
    int64_t ldc16 = 0;
    int relu = 0;
    const float* R32 = nullptr;
    int64_t ldr32 = 0;
    const uint16_t* R16 = nullptr;
    int64_t ldr16 = 0;
    const float* ln_g = nullptr;
    const float* ln_b = nullptr;
    // 3-product GEMM with fp32-class accuracy: A16 holds [hi | hi | lo] rows of 3 K columns (split3_f16_rows), the
    // cached weight copy is [hi | lo | hi]; K stays the fp32 K
    int split3 = 0;
};
int tc_gemm_run(TcWeights* t, const TcGemmArgs& g, cudaStream_t s);

// y[rows, 3 K] (fp16, dense) = [hi | hi | lo] with hi = fp16(x), lo = fp16(x - hi)
int split3_f16_rows(const float* x, int64_t ldx, uint16_t* y, int64_t rows, int K, cudaStream_t s);
// y[rows, cols] (fp16, dense) = saturating round-to-nearest of x[rows, ldx] (cols % 4 == 0)
int f32_to_f16_rows(const float* x, int64_t ldx, uint16_t* y, int64_t rows, int cols, cudaStream_t s);

}  // namespace cone

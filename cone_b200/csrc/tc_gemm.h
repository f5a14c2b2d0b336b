// Tensor-core (tcgen05 + TMA + TMEM) GEMM path, CONE_PREC_TC.  See tc_gemm.cu.
#pragma once
#include "common.cuh"

namespace cone {

struct TcWeights;  // bf16 weight copies + TMA descriptors, owned by the cone_weights handle

int tc_weights_create(TcWeights** out, cudaStream_t s);
void tc_weights_destroy(TcWeights* t);
// per-call activation staging (bf16 copy of A) lives in the caller's workspace
size_t tc_scratch_bytes(int64_t max_rows, int max_k);
void tc_set_scratch(TcWeights* t, void* scratch, size_t bytes);
bool tc_gemm_supported(int64_t M, int N, int K);
// y[M,N] = epi(x[M,K] * W[N,K]^T + b (+R)), bf16 operands, fp32 accumulate, fp32 output
int tc_gemm(TcWeights* t, const float* x, int64_t ldx, int64_t M, const float* W, const float* b, int N, int K, float* y,
            int64_t ldy, int relu, const float* R, int64_t ldr, cudaStream_t s);

// fp16 operand already in memory (row pitch lda halves); optional fp32 (C32) and fp16 (C16) outputs; optional fused
// residual (R fp32) and LayerNorm over the row (ln_g/ln_b non-null requires N == 256).  fp16 pointers as uint16_t*.
int tc_gemm_f16(TcWeights* t, const uint16_t* A16, int64_t lda, int64_t M, const float* W, const float* b, int N, int K,
                float* C32, int64_t ldc32, uint16_t* C16, int64_t ldc16, int relu, const float* R, int64_t ldr,
                const float* ln_g, const float* ln_b, cudaStream_t s);

}  // namespace cone

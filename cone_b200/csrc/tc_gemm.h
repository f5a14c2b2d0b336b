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

}  // namespace cone

// Tensor-core (tcgen05 + TMA + TMEM) GEMM path, CONE_PREC_TC.  See tc_gemm.cu.
#pragma once
#include <cuda.h>

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
    // 2-product GEMM for an operand that only exists in fp16 (the attention-pooled memory): A16 holds K columns and is
    // read twice against the first two thirds [hi | lo] of the split weight copy — the weights keep fp32-class accuracy
    int wsplit = 0;
};
int tc_gemm_run(TcWeights* t, const TcGemmArgs& g, cudaStream_t s);

// ---- shared with enc_tail.cu
// fp16 copy [N, K] of an fp32 nn.Linear weight, cached in (and owned by) the handle
int tc_weight_f16(TcWeights* t, const float* W, int N, int K, cudaStream_t s, const uint16_t** out);
// 2-D tensor map of a row-major [rows, cols] tensor with row pitch ld (elements); box = [box_cols x box_rows] with a
// 128-byte inner extent, 128-byte swizzle
int tc_make_map(CUtensorMap* map, const void* base, bool f32, int64_t rows, int64_t cols, int64_t ld, int box_cols,
                int box_rows);
int tc_num_sms(const TcWeights* t);

// Fused encoder-layer tail (enc_tail.cu): out = LN2(x + FFN(x)), x = LN1(res + att . Wo^T + bo), rows of d = 256.
// The residual stream is fp16 hi (+ optional fp16 lo = value - hi); out_hi / out_lo may alias res_hi / res_lo.
struct EncTailArgs {
    const uint16_t* att16 = nullptr;   // [M, d] attention output
    int64_t lda = 0;
    const uint16_t* res_hi = nullptr;  // [M, d]
    const uint16_t* res_lo = nullptr;  // nullable
    int64_t ldr = 0;
    uint16_t* out_hi = nullptr;
    uint16_t* out_lo = nullptr;        // nullable
    int64_t ldo = 0;
    float* C32 = nullptr;              // nullable fp32 copy of the output
    int64_t ldc32 = 0;
    int64_t M = 0;
    int d = 0, ffn = 0;
    const float *Wo = nullptr, *bo = nullptr, *ln1_g = nullptr, *ln1_b = nullptr;
    const float *W1 = nullptr, *b1 = nullptr, *W2 = nullptr, *b2 = nullptr, *ln2_g = nullptr, *ln2_b = nullptr;
    int cta_group = 0;                 // 1 or 2; 0 = default (2, or env CONE_ENC_TAIL_CG)
    // gather mode (g_vid_base != nullptr): res_hi / res_lo are the SOURCE tables [g_nsrc, d] = per-frame rows [g_nvid] followed by
    // the per-token rows, and residual row m = window b = m / g_S, row r = m % g_S is fetched from source row
    // min(g_vid_base[b] + r, g_nvid - 1) (r < g_Lv) or g_nvid + g_txt_base[b] + r - g_Lv by TMA gather4 (ldr == d)
    const int64_t* g_vid_base = nullptr;
    const int64_t* g_txt_base = nullptr;
    int g_S = 0, g_Lv = 0;
    int64_t g_nvid = 0, g_nsrc = 0;
};
int enc_tail_supported(int d, int ffn);
int enc_tail_run(TcWeights* t, const EncTailArgs& a, cudaStream_t s);

// y[rows, 3 K] (fp16, dense) = [hi | hi | lo] with hi = fp16(x), lo = fp16(x - hi)
int split3_f16_rows(const float* x, int64_t ldx, uint16_t* y, int64_t rows, int K, cudaStream_t s);
// y[rows, cols] (fp16, dense) = saturating round-to-nearest of x[rows, ldx] (cols % 4 == 0)
int f32_to_f16_rows(const float* x, int64_t ldx, uint16_t* y, int64_t rows, int cols, cudaStream_t s);

}  // namespace cone

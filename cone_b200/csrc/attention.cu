// Small-sequence multi-head attention for Moment-DETR windows (S = Lv + Lt <= 256 rows, head_dim 32).
// Math follows torch.nn.functional.multi_head_attention_forward as called at cone/transformer.py:239,
// 304, 308: q scaled by sqrt(1/head_dim) BEFORE the dot product, additive -inf key-padding mask,
// softmax over keys, then P.V.  One CTA per (window, head); K and V of the head live in shared memory.
#include <math_constants.h>

#include "kernels.h"

namespace cone {

namespace {

constexpr int HD = 32;            // head dim (hidden 256 / 8 heads)
constexpr int MAX_S = 256;        // max keys per window
constexpr int ENC_WARPS = 4;

__device__ __forceinline__ bool key_valid(int j, int Lv, int vl, int tl) { return j < Lv ? (j < vl) : ((j - Lv) < tl); }

__global__ void __launch_bounds__(ENC_WARPS * 32)
enc_self_attention_kernel(const float* __restrict__ qk, int64_t ldqk, const float* __restrict__ v, int64_t ldv,
                          float* __restrict__ o, int64_t ldo, const int32_t* __restrict__ vlen,
                          const int32_t* __restrict__ tlen, int Lv, int Lt, int d_model) {
    extern __shared__ float smem[];
    const int S = Lv + Lt;
    float* Ks = smem;                        // [S][HD+1]
    float* Vs = Ks + S * (HD + 1);           // [S][HD]
    float* Qs = Vs + S * HD;                 // [ENC_WARPS][HD]
    float* Ps = Qs + ENC_WARPS * HD;         // [ENC_WARPS][S]
    const int64_t b = blockIdx.x;
    const int h = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int vl = vlen[b], tl = tlen[b];
    const int64_t row0 = b * S;
    const float scale = 0.17677669529663687f;  // sqrt(1/32), as torch scales q

    for (int i = threadIdx.x; i < S * HD; i += blockDim.x) {
        const int j = i / HD, c = i % HD;
        Ks[j * (HD + 1) + c] = qk[(row0 + j) * ldqk + d_model + h * HD + c];
        Vs[j * HD + c] = v[(row0 + j) * ldv + h * HD + c];
    }
    __syncthreads();

    constexpr int KPL = MAX_S / 32;  // keys per lane
    for (int i = warp; i < S; i += ENC_WARPS) {
        Qs[warp * HD + lane] = qk[(row0 + i) * ldqk + h * HD + lane] * scale;
        __syncwarp();
        float sc[KPL];
        float mx = -CUDART_INF_F;
#pragma unroll
        for (int t = 0; t < KPL; ++t) {
            const int j = lane + 32 * t;
            float a = -CUDART_INF_F;
            if (j < S && key_valid(j, Lv, vl, tl)) {
                a = 0.f;
#pragma unroll
                for (int c = 0; c < HD; ++c) a = fmaf(Qs[warp * HD + c], Ks[j * (HD + 1) + c], a);
            }
            sc[t] = a;
            mx = fmaxf(mx, a);
        }
        mx = warp_max(mx);
        float sum = 0.f;
#pragma unroll
        for (int t = 0; t < KPL; ++t) {
            const int j = lane + 32 * t;
            const float e = (sc[t] == -CUDART_INF_F) ? 0.f : expf(sc[t] - mx);
            if (j < S) Ps[warp * S + j] = e;
            sum += e;
        }
        sum = warp_sum(sum);
        __syncwarp();
        float acc = 0.f;
        for (int j = 0; j < S; ++j) acc = fmaf(Ps[warp * S + j], Vs[j * HD + lane], acc);
        // all keys masked (cannot happen for real windows) -> 0/0 = NaN, like torch's softmax of all -inf
        o[(row0 + i) * ldo + h * HD + lane] = acc / sum;
        __syncwarp();
    }
}

// decoder self-attention: nq <= 8 slots, one warp per (window, head), lane = channel
__global__ void dec_self_attention_kernel(const float* __restrict__ qk, int64_t ldqk, const float* __restrict__ v,
                                          int64_t ldv, float* __restrict__ o, int64_t ldo, int64_t B, int nq,
                                          int nheads, int d_model) {
    const int lane = threadIdx.x & 31;
    const int64_t wid = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (wid >= B * nheads) return;
    const int64_t b = wid / nheads;
    const int h = (int)(wid % nheads);
    const float scale = 0.17677669529663687f;
    float q[8], k[8], vv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        if (i < nq) {
            const int64_t row = b * nq + i;
            q[i] = qk[row * ldqk + h * HD + lane] * scale;
            k[i] = qk[row * ldqk + d_model + h * HD + lane];
            vv[i] = v[row * ldv + h * HD + lane];
        } else {
            q[i] = k[i] = vv[i] = 0.f;
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        if (i >= nq) break;
        float sc[8];
        float mx = -CUDART_INF_F;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (j < nq) {
                sc[j] = warp_sum(q[i] * k[j]);
                mx = fmaxf(mx, sc[j]);
            }
        }
        float sum = 0.f, acc = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (j < nq) {
                const float e = expf(sc[j] - mx);
                sum += e;
                acc = fmaf(e, vv[j], acc);
            }
        }
        o[(b * nq + i) * ldo + h * HD + lane] = acc / sum;
    }
}

// decoder cross-attention: one warp per (window, head); nq <= 8 queries against S memory keys
__global__ void __launch_bounds__(128)
dec_cross_attention_kernel(const float* __restrict__ q, int64_t ldq, const float* __restrict__ k, int64_t ldk,
                           const float* __restrict__ v, int64_t ldv, float* __restrict__ o, int64_t ldo,
                           const int32_t* __restrict__ vlen, const int32_t* __restrict__ tlen, int64_t B, int nq, int Lv,
                           int Lt, int nheads) {
    extern __shared__ float smem[];
    const int S = Lv + Lt;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t wid = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    if (wid >= B * nheads) return;  // whole warp exits together
    float* Qs = smem + warp * (8 * HD + 8 * S);  // [8][HD]
    float* Ps = Qs + 8 * HD;                     // [8][S]
    const int64_t b = wid / nheads;
    const int h = (int)(wid % nheads);
    const int vl = vlen[b], tl = tlen[b];
    const float scale = 0.17677669529663687f;
    for (int i = 0; i < nq; ++i) Qs[i * HD + lane] = q[(b * nq + i) * ldq + h * HD + lane] * scale;
    __syncwarp();
    float mx[8], sum[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { mx[i] = -CUDART_INF_F; sum[i] = 0.f; }
    // pass 1: scores (each lane owns keys lane, lane+32, ...)
    for (int j = lane; j < S; j += 32) {
        const bool ok = key_valid(j, Lv, vl, tl);
        float kr[HD];
        const float4* kp = reinterpret_cast<const float4*>(k + (b * S + j) * ldk + h * HD);
#pragma unroll
        for (int c = 0; c < HD / 4; ++c) {
            const float4 t = kp[c];
            kr[4 * c] = t.x; kr[4 * c + 1] = t.y; kr[4 * c + 2] = t.z; kr[4 * c + 3] = t.w;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (i < nq) {
                float a = -CUDART_INF_F;
                if (ok) {
                    a = 0.f;
#pragma unroll
                    for (int c = 0; c < HD; ++c) a = fmaf(Qs[i * HD + c], kr[c], a);
                }
                Ps[i * S + j] = a;
                mx[i] = fmaxf(mx[i], a);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) mx[i] = warp_max(mx[i]);
    __syncwarp();
    for (int j = lane; j < S; j += 32) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (i < nq) {
                const float a = Ps[i * S + j];
                const float e = (a == -CUDART_INF_F) ? 0.f : expf(a - mx[i]);
                Ps[i * S + j] = e;
                sum[i] += e;
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) sum[i] = warp_sum(sum[i]);
    __syncwarp();
    // pass 2: P.V, lane = channel
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    for (int j = 0; j < S; ++j) {
        const float vj = v[(b * S + j) * ldv + h * HD + lane];
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (i < nq) acc[i] = fmaf(Ps[i * S + j], vj, acc[i]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
        if (i < nq) o[(b * nq + i) * ldo + h * HD + lane] = acc[i] / sum[i];
}

}  // namespace

int enc_self_attention(const float* qk, int64_t ldqk, const float* v, int64_t ldv, float* o, int64_t ldo,
                       const int32_t* vlen, const int32_t* tlen, int64_t B, int Lv, int Lt, int nheads,
                       cudaStream_t s) {
    if (B == 0) return CONE_OK;
    const int S = Lv + Lt;
    CONE_REQUIRE(S <= MAX_S, "enc_self_attention: window of %d rows exceeds %d", S, MAX_S);
    const size_t smem = sizeof(float) * ((size_t)S * (HD + 1) + (size_t)S * HD + ENC_WARPS * HD + (size_t)ENC_WARPS * S);
    static bool attr_set = false;
    if (!attr_set) {
        CONE_CUDA(cudaFuncSetAttribute(enc_self_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        attr_set = true;
    }
    dim3 grid((unsigned)B, (unsigned)nheads);
    ProfScope ps(s, P_ENC_ATTN, 4.0 * (double)B * nheads * S * S * HD, 16.0 * (double)B * S * nheads * HD);
    enc_self_attention_kernel<<<grid, ENC_WARPS * 32, smem, s>>>(qk, ldqk, v, ldv, o, ldo, vlen, tlen, Lv, Lt,
                                                               nheads * HD);
    CONE_LAUNCH_CHECK("enc_self_attention");
    return CONE_OK;
}

int dec_self_attention(const float* qk, int64_t ldqk, const float* v, int64_t ldv, float* o, int64_t ldo, int64_t B,
                       int nq, int nheads, cudaStream_t s) {
    if (B == 0) return CONE_OK;
    CONE_REQUIRE(nq <= 8, "dec_self_attention: at most 8 moment slots");
    const int warps = 4;
    ProfScope ps(s, P_DEC_ATTN);
    dec_self_attention_kernel<<<(unsigned)cdiv64(B * nheads, warps), warps * 32, 0, s>>>(qk, ldqk, v, ldv, o, ldo, B, nq,
                                                                                       nheads, nheads * HD);
    CONE_LAUNCH_CHECK("dec_self_attention");
    return CONE_OK;
}

int dec_cross_attention(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv,
                        float* o, int64_t ldo, const int32_t* vlen, const int32_t* tlen, int64_t B, int nq, int Lv,
                        int Lt, int nheads, cudaStream_t s) {
    if (B == 0) return CONE_OK;
    const int S = Lv + Lt;
    CONE_REQUIRE(nq <= 8 && S <= MAX_S, "dec_cross_attention: unsupported nq=%d S=%d", nq, S);
    CONE_REQUIRE((ldk & 3) == 0, "dec_cross_attention: ldk must be a multiple of 4");
    const int warps = 4;
    const size_t smem = sizeof(float) * warps * (8 * HD + 8 * (size_t)S);
    ProfScope ps(s, P_DEC_ATTN, 4.0 * (double)B * nheads * nq * S * HD, 8.0 * (double)B * S * nheads * HD);
    dec_cross_attention_kernel<<<(unsigned)cdiv64(B * nheads, warps), warps * 32, smem, s>>>(
        q, ldq, k, ldk, v, ldv, o, ldo, vlen, tlen, B, nq, Lv, Lt, nheads);
    CONE_LAUNCH_CHECK("dec_cross_attention");
    return CONE_OK;
}

}  // namespace cone

// Tensor-core GEMM for the dense projections (CONE_PREC_TC): C = epi(A * W^T), fp16 operands, fp32 accumulate.
//
// sm_100a design: persistent CTAs (one per SM), warp-specialised:
//   warp 0    TMA producer  cp.async.bulk.tensor 2-D boxes of A [128 x 64] and W [BN x 64] (fp16, 128-byte swizzle)
//                           into a 3-stage shared-memory ring, completion on mbarriers
//   warp 1    MMA issuer    one thread issues tcgen05.mma (cta_group::1, kind::f16, M=128, N=BN, K=16) from
//                           shared-memory descriptors; accumulators live in TMEM (2 x BN fp32 columns, double
//                           buffered so the epilogue of tile i overlaps the MMAs of tile i+1)
//   warps 2-9 epilogue      tcgen05.ld TMEM -> registers (one accumulator row per thread, two warps per TMEM lane
//                           quarter splitting the columns: 2 epilogue warps per scheduler hide each other's
//                           latencies); bias / fp16 residual /
//                           ReLU / LayerNorm(N = 256) in registers; results are written as 16-byte vectors into a
//                           128-byte-swizzled shared-memory box and leave the SM as ONE TMA store per 32 x 64 box
//                           (no per-row store instructions, out-of-range rows clipped by the tensor map); the
//                           residual tile arrives the same way through a TMA load.
// All GEMMs of this model have K = 256 or 1024 and N <= 1024: they are bound by HBM traffic of activations, not
// by the tensor pipe, so the epilogue's job is to keep every byte coalesced and the instruction count low.
// fp16 (11 significant bits) rather than bf16: the reference comparison needs 1e-3 on spans and scores, which
// bf16 operands miss by 3-5x (measured by emulation); conversions saturate.
#include <cuda.h>
#include <stdlib.h>
#include <cuda_fp16.h>

#include <map>
#include <vector>

#include "kernels.h"
#include "tc_gemm.h"
#include "tc_ptx.cuh"

namespace cone {

using namespace ptx;

namespace {

constexpr int BM = 128;          // UMMA M
constexpr int BK = 64;           // fp16 elements per stage row = 128 bytes = one swizzle atom row
constexpr int UMMA_K = 16;
constexpr int STAGES = 3;
constexpr int A_BYTES = BM * BK * 2;  // 16 KB
constexpr int TC_THREADS = 320;  // producer warp + MMA warp + 8 epilogue warps
constexpr int BOX_BYTES = 32 * 128;   // one epilogue box: 32 rows x 128 bytes (64 fp16 or 32 fp32 columns)

struct TcEpilogue {
    const float* bias;   // [N] or null
    const float* R32;    // fp32 residual [M, ldr32] or null (small-M decoder GEMMs)
    int64_t ldr32;
    int has_r16;         // fp16 residual through tensor map tmR
    int has_c16;         // fp16 output through tensor map tmC16
    int has_c32;         // fp32 output through tensor map tmC32
    int relu;
    const float* ln_g;   // fused LayerNorm over the N = BN columns of the row (null = off)
    const float* ln_b;
    float ln_eps;
};

__device__ __forceinline__ void add_bias64(float* x, const float* bias) {
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(bias) + q);
        x[4 * q] += b.x; x[4 * q + 1] += b.y; x[4 * q + 2] += b.z; x[4 * q + 3] += b.w;
    }
}

// ------------------------------------------------------------------------------------------ kernel
constexpr int EPI_WARPS = 8;  // two warps per TMEM lane quarter; each takes half of the tile's columns

// WRES ("weights resident", K = 4 k-blocks = 256): the CTA is bound to ONE n-tile, its whole [BN x 256] fp16 weight
// tile (128 KB) is loaded once and stays in shared memory, and the ring carries only A (16 KB per k-block, 4 stages).
// Why: ncu showed every K = 256 GEMM of the model stuck at ~1450 cycles per k-block against 512 cycles of MMA work,
// whatever the epilogue did: with A + W in a 3-stage ring only 144 KB per SM were in flight against ~2 us of loaded
// memory latency, and two thirds of those bytes were the same weight tile streamed again for every M tile.
constexpr int WRES_KBLOCKS = 4;
constexpr int WRES_STAGES = 4;
constexpr int FAST_STAGES = 4;  // A + W ring depth of the streamlined epilogues when the weights are not resident

// Epilogue variants.  The shape that carries most of the model's 2.9 M rows is compile-time specialised (no flag
// tests, 64-column steps, accumulator released right after the second TMEM load, one staging box per warp):
//   EPI_PLAIN16: bias (+ReLU) -> fp16                                         (QKV, FFN1, decoder q | Wk^T q)
//   EPI_LN16:    the generic code below with its flags fixed at compile time to bias + fp16 residual (TMA boxes) ->
//                LayerNorm -> fp16                                             (out_proj, FFN2)
//   EPI_GENERIC: every flag at run time: residual through TMA boxes, LayerNorm over the 256-wide row (out_proj,
//                FFN2), fp32 in/out decoder-side GEMMs, cone_linear, BN = 128
// (A specialised LayerNorm epilogue that read the residual with per-row global loads and kept the 128 columns of a
// row in registers was measured 20-50 % SLOWER than the generic TMA-box one and was dropped: profiles/r01_notes.md.)
// ncu on the generic epilogue at 2 warps per scheduler: ~9 issue cycles per instruction, a third of the stalls in
// LDCU -> UISETP -> BRA chains of the run-time flags, another tenth on bias loads issued after the TMEM wait.
enum { EPI_PLAIN16 = 0, EPI_LN16 = 1, EPI_GENERIC = 2 };

// 64 fp32 values of one row -> fp16 -> this lane's row of a [32 x 64] 128-byte-swizzled box -> one TMA store
__device__ __forceinline__ void store_box16(const float* x, uint8_t* box, const CUtensorMap* map, int col, int row0,
                                            int lane) {
    if (lane == 0) tma_store_wait_read<0>();  // the previous store has finished reading the box
    __syncwarp();
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        uint4 v;
        v.x = pack_h2(x[8 * u], x[8 * u + 1]);
        v.y = pack_h2(x[8 * u + 2], x[8 * u + 3]);
        v.z = pack_h2(x[8 * u + 4], x[8 * u + 5]);
        v.w = pack_h2(x[8 * u + 6], x[8 * u + 7]);
        *reinterpret_cast<uint4*>(box + sw128(lane, u)) = v;
    }
    fence_async_smem();
    __syncwarp();
    if (lane == 0) {
        tma_store_2d(map, box, col, row0);
        tma_store_commit();
    }
}

template <int BN, bool WRES, int MODE>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmC16,
               const __grid_constant__ CUtensorMap tmC32, TcEpilogue ep, int64_t M, int N, int K, int a_wrap) {
    constexpr int B_BYTES = BN * BK * 2;
    constexpr int HALF = BN / 2;  // columns per epilogue warp
    constexpr bool FAST = MODE == EPI_PLAIN16;
    constexpr int NSTAGE = WRES ? WRES_STAGES : (FAST ? FAST_STAGES : STAGES);
    constexpr int B_RING = WRES ? WRES_KBLOCKS * B_BYTES : NSTAGE * B_BYTES;  // resident tile or ring
    constexpr int N_BOX = (WRES || FAST) ? EPI_WARPS : 2 * EPI_WARPS;         // one box per warp unless generic + ring
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem;
    if (WRES || FAST) {  // no room for alignment slack: the dynamic window must already be 1 KB aligned
        smem = smem_raw;
        if ((smem_u32(smem_raw) & 1023u) != 0) __trap();
    } else {
        smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    }
    uint8_t* sA = smem;
    uint8_t* sB = sA + NSTAGE * A_BYTES;
    uint8_t* sOut = sB + B_RING;                     // one TMA-store box per epilogue warp
    uint8_t* sRes = WRES ? sOut : sOut + EPI_WARPS * BOX_BYTES;  // one TMA-loaded residual box per epilogue warp
    uint64_t* full = reinterpret_cast<uint64_t*>(sOut + N_BOX * BOX_BYTES);
    uint64_t* empty = full + NSTAGE;
    uint64_t* tfull = empty + NSTAGE;
    uint64_t* tempty = tfull + 2;
    uint64_t* rfull = tempty + 2;  // one per epilogue warp
    uint64_t* wfull = rfull + EPI_WARPS;
    float2* ln_part = reinterpret_cast<float2*>(wfull + 1);  // [EPI_WARPS][32] partial (sum, sum of squares)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ln_part + EPI_WARPS * 32);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t m_tiles = (M + BM - 1) / BM;
    const int n_tiles = N / BN;
    const int64_t tiles = m_tiles * n_tiles;
    const int num_k = K / BK;
    // tile walk: WRES -> this CTA's n-tile is fixed and t runs over M tiles; otherwise t runs over all (m, n) tiles
    const int n_fixed = WRES ? (int)(blockIdx.x % n_tiles) : 0;
    const int64_t t_begin = WRES ? (int64_t)(blockIdx.x / n_tiles) : (int64_t)blockIdx.x;
    const int64_t t_step = WRES ? (int64_t)(gridDim.x / n_tiles) : (int64_t)gridDim.x;
    const int64_t t_end = WRES ? m_tiles : tiles;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
        for (int i = 0; i < NSTAGE; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        mbar_init(wfull, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull[i], 1);
            mbar_init(&tempty[i], FAST ? EPI_WARPS : EPI_WARPS * 32);
        }
        for (int i = 0; i < EPI_WARPS; ++i) mbar_init(&rfull[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // TMEM: 2 accumulators of BN fp32 columns (512 columns = the whole TMEM for BN = 256)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)(2 * BN))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {  // ---------------------------------------------------------------- TMA producer
            int stage = 0;
            uint32_t phase = 0;
            if (WRES) {  // the weight tile of this CTA's n-tile, once
                mbar_expect_tx(wfull, WRES_KBLOCKS * B_BYTES);
                for (int kb = 0; kb < WRES_KBLOCKS; ++kb) tma_load_2d(sB + kb * B_BYTES, &tmB, wfull, kb * BK, n_fixed * BN);
            }
            for (int64_t t = t_begin; t < t_end; t += t_step) {
                const int m0 = (int)(WRES ? t : t / n_tiles) * BM, n0 = (WRES ? n_fixed : (int)(t % n_tiles)) * BN;
                for (int kb = 0; kb < num_k; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    mbar_expect_tx(&full[stage], WRES ? A_BYTES : A_BYTES + B_BYTES);
                    tma_load_2d(sA + stage * A_BYTES, &tmA, &full[stage], (kb * BK) % a_wrap, m0);  // a_wrap < K: A read twice
                    if (!WRES) tma_load_2d(sB + stage * B_BYTES, &tmB, &full[stage], kb * BK, n0);
                    if (++stage == NSTAGE) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        {  // ---------------------------------------------------------------------------------- MMA issuer
            // The whole warp runs the role (uniform control flow); one elected lane issues tcgen05.mma / commit, so that
            // descriptors and addresses stay in uniform registers (a lane-0-only role made the compiler wrap every MMA
            // in an ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall: ~120 issue cycles per MMA, tc_ptx.cuh).
            // instruction descriptor: D fp32, A/B fp16 K-major, N = BN, M = 128
            constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            const uint32_t a_lo0 = desc_lo_sw128(smem_u32(sA)), b_lo0 = desc_lo_sw128(smem_u32(sB));
            const uint32_t tmem_base_u = __shfl_sync(0xffffffffu, tmem_base, 0);  // warp-uniform for the compiler
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            if (WRES) mbar_wait(wfull, 0);
            for (int64_t t = t_begin; t < t_end; t += t_step) {
                mbar_wait(&tempty[acc], acc_phase ^ 1);  // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base_u + (uint32_t)(acc * BN);
                for (int kb = 0; kb < num_k; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t a_lo = a_lo0 + (uint32_t)stage * (A_BYTES >> 4);
                    const uint32_t b_lo = b_lo0 + (uint32_t)(WRES ? kb : stage) * (B_BYTES >> 4);
                    if (elect_one_sync()) {
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k)  // +32 bytes per K = 16 step = +2 in the start-address field
                            umma_f16_lo(d_tmem, a_lo + 2 * k, b_lo + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                        umma_commit(&empty[stage]);  // frees the smem stage once these MMAs have read it
                        if (kb == num_k - 1) umma_commit(&tfull[acc]);  // accumulator complete -> epilogue
                    }
                    __syncwarp();
                    if (++stage == NSTAGE) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                if (++acc == 2) {
                    acc = 0;
                    acc_phase ^= 1;
                }
            }
        }
    } else if (FAST) {  // ------------------------------------------------- epilogue, specialised (see enum above)
        const int ew = warp - 2;       // 0..7
        const int quarter = warp & 3;  // TMEM lane quarter this warp may read
        const int half = ew >> 2;      // which half of the tile's columns
        uint8_t* box = sOut + ew * BOX_BYTES;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int64_t t = t_begin; t < t_end; t += t_step) {
            const int64_t m0 = (WRES ? t : t / n_tiles) * BM;
            const int n0 = (WRES ? n_fixed : (int)(t % n_tiles)) * BN + half * HALF;  // first of this warp's 128 columns
            const int row0 = (int)m0 + quarter * 32;                                  // first of this warp's 32 rows
            mbar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN + half * HALF);
            auto release_acc = [&]() {  // this warp's slice of the accumulator is in registers
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[acc]);
                if (++acc == 2) {
                    acc = 0;
                    acc_phase ^= 1;
                }
            };
            auto add_bias = [&](float* x, int c) {
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    const float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + n0 + c) + q);
                    x[4 * q] += b.x; x[4 * q + 1] += b.y; x[4 * q + 2] += b.z; x[4 * q + 3] += b.w;
                }
            };
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                float x[64];
                tmem_ld_32x64(taddr + ch * 64, x);
                tmem_ld_wait();
                if (ch == 1) release_acc();
                add_bias(x, ch * 64);
                if (ep.relu) {
#pragma unroll
                    for (int j = 0; j < 64; ++j) x[j] = fmaxf(x[j], 0.f);
                }
                store_box16(x, box, &tmC16, n0 + ch * 64, row0, lane);
            }
        }
        if (lane == 0) tma_store_wait_read<0>();  // shared memory must outlive the last bulk stores
    } else {  // ------------------------------------------------------------------- epilogue, generic (run-time flags)
        const int ew = warp - 2;       // 0..7
        const int quarter = warp & 3;  // TMEM lane quarter this warp may read
        const int half = ew >> 2;      // which half of the tile's columns
        const int partner = ew ^ 4;    // the warp with the same rows and the other columns
        uint8_t* out_box = sOut + ew * BOX_BYTES;
        uint8_t* res_box = sRes + ew * BOX_BYTES;
        uint64_t* rbar = &rfull[ew];
        uint32_t rphase = 0;
        // EPI_LN16 fixes the flags at compile time (the run-time tests were LDCU -> UISETP -> BRA chains in every chunk)
        constexpr bool LNM = MODE == EPI_LN16;
        const bool ln = LNM ? true : ep.ln_g != nullptr;
        const bool f_r16 = LNM ? true : ep.has_r16 != 0;
        const bool f_c16 = LNM ? true : ep.has_c16 != 0;
        const bool f_c32 = LNM ? false : ep.has_c32 != 0;
        const bool f_bias = LNM ? true : ep.bias != nullptr;
        const float* f_R32 = LNM ? nullptr : ep.R32;
        const bool f_relu = LNM ? false : ep.relu != 0;
        int acc = 0;
        uint32_t acc_phase = 0;

        for (int64_t t = t_begin; t < t_end; t += t_step) {
            const int64_t m0 = (WRES ? t : t / n_tiles) * BM;
            const int n0 = (WRES ? n_fixed : (int)(t % n_tiles)) * BN + half * HALF;  // first column of this warp's half
            const int row0 = (int)m0 + quarter * 32;               // first row of this warp's 32 rows
            mbar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN + half * HALF);

            // x[0..31] = acc + bias (+ residual) for columns [c, c+32) of this warp's half, one row per lane.
            // The fp16 residual arrives as a [32 x 64] box: requested at even 32-column steps, consumed in two halves.
            auto load_sub = [&](int c, float* x) {
                const int sub = (c >> 5) & 1;
                if (f_r16 && sub == 0 && lane == 0) {
                    // res_box doubles as the fp32 store box (see below) and, in WRES mode, as the result box
                    if (WRES || f_c32) tma_store_wait_read<0>();
                    mbar_expect_tx(rbar, BOX_BYTES);
                    tma_load_2d(res_box, &tmR, rbar, n0 + c, row0);
                }
                tmem_ld_32x32(taddr + c, x);
                tmem_ld_wait();
                if (f_bias) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + n0 + c) + q);
                        x[4 * q] += b.x; x[4 * q + 1] += b.y; x[4 * q + 2] += b.z; x[4 * q + 3] += b.w;
                    }
                }
                if (f_r16) {
                    if (sub == 0) {
                        mbar_wait(rbar, rphase);
                        rphase ^= 1;
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const uint4 v = *reinterpret_cast<const uint4*>(res_box + sw128(lane, sub * 4 + u));
                        const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float2 f = __half22float2(h[e]);
                            x[8 * u + 2 * e] += f.x;
                            x[8 * u + 2 * e + 1] += f.y;
                        }
                    }
                    if (sub == 1) __syncwarp();  // every lane has read the box before the next TMA load overwrites it
                }
                if (f_R32) {  // small-M path: plain row loads
                    const int64_t row = (int64_t)row0 + lane;
                    if (row < M) {
                        const float4* r4 = reinterpret_cast<const float4*>(f_R32 + row * ep.ldr32 + n0 + c);
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const float4 b = r4[q];
                            x[4 * q] += b.x; x[4 * q + 1] += b.y; x[4 * q + 2] += b.z; x[4 * q + 3] += b.w;
                        }
                    }
                }
            };

            float mean = 0.f, rstd = 1.f;
            if (ln) {  // pass 1: pre-LayerNorm values back into TMEM + row statistics (two warps per row)
                float s1 = 0.f, s2 = 0.f;
                for (int c = 0; c < HALF; c += 32) {
                    float x[32];
                    load_sub(c, x);
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        s1 += x[j];
                        s2 = fmaf(x[j], x[j], s2);
                    }
                    tmem_st_32x32(taddr + c, x);
                    tmem_st_wait();
                }
                ln_part[ew * 32 + lane] = make_float2(s1, s2);
                asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");  // the two warps of this row quarter
                const float2 o = ln_part[partner * 32 + lane];
                asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");  // partials consumed: slot reusable
                s1 += o.x;
                s2 += o.y;
                mean = s1 * (1.f / BN);
                rstd = rsqrtf(fmaxf(s2 * (1.f / BN) - mean * mean, 0.f) + ep.ln_eps);
            }
            for (int c = 0; c < HALF; c += 32) {
                float x[32];
                if (ln) {
                    tmem_ld_32x32(taddr + c, x);
                    tmem_ld_wait();
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 g = __ldg(reinterpret_cast<const float4*>(ep.ln_g + n0 + c) + q);
                        const float4 b = __ldg(reinterpret_cast<const float4*>(ep.ln_b + n0 + c) + q);
                        x[4 * q] = (x[4 * q] - mean) * rstd * g.x + b.x;
                        x[4 * q + 1] = (x[4 * q + 1] - mean) * rstd * g.y + b.y;
                        x[4 * q + 2] = (x[4 * q + 2] - mean) * rstd * g.z + b.z;
                        x[4 * q + 3] = (x[4 * q + 3] - mean) * rstd * g.w + b.w;
                    }
                } else {
                    load_sub(c, x);
                    if (f_relu) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) x[j] = fmaxf(x[j], 0.f);
                    }
                }
                if (f_c16) {  // two 32-column steps fill one [32 rows x 64 fp16] 128-byte-swizzled box
                    const int sub = (c >> 5) & 1;
                    if (sub == 0) {
                        if (lane == 0) tma_store_wait_read<0>();  // previous store has released the box
                        __syncwarp();
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        __half2 h0 = __floats2half2_rn(x[8 * u], x[8 * u + 1]);
                        __half2 h1 = __floats2half2_rn(x[8 * u + 2], x[8 * u + 3]);
                        __half2 h2 = __floats2half2_rn(x[8 * u + 4], x[8 * u + 5]);
                        __half2 h3 = __floats2half2_rn(x[8 * u + 6], x[8 * u + 7]);
                        uint4 v;
                        v.x = *reinterpret_cast<uint32_t*>(&h0);
                        v.y = *reinterpret_cast<uint32_t*>(&h1);
                        v.z = *reinterpret_cast<uint32_t*>(&h2);
                        v.w = *reinterpret_cast<uint32_t*>(&h3);
                        *reinterpret_cast<uint4*>(out_box + sw128(lane, sub * 4 + u)) = v;
                    }
                    if (sub == 1) {
                        fence_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            tma_store_2d(&tmC16, out_box, n0 + c - 32, row0);
                            tma_store_commit();
                        }
                    }
                }
                if (f_c32) {  // one box of 32 rows x 32 fp32 columns
                    // with an fp16 output in flight the fp32 box is staged in res_box (free in this pass: the host
                    // side only allows both outputs together with LayerNorm or without an fp16 residual)
                    uint8_t* box32 = f_c16 ? res_box : out_box;
                    if (lane == 0) tma_store_wait_read<0>();
                    __syncwarp();
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        *reinterpret_cast<float4*>(box32 + sw128(lane, u)) =
                            make_float4(x[4 * u], x[4 * u + 1], x[4 * u + 2], x[4 * u + 3]);
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&tmC32, box32, n0 + c, row0);
                        tma_store_commit();
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&tempty[acc]);
            if (++acc == 2) {
                acc = 0;
                acc_phase ^= 1;
            }
        }
        if (lane == 0) tma_store_wait_read<0>();  // shared memory must outlive the last bulk stores
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(2 * BN))
                     : "memory");
    }
}

// fp32 -> fp16 (round to nearest, saturating) for GEMM operands
__global__ void f32_to_f16_kernel(const float* __restrict__ x, int64_t ldx, __half* __restrict__ y, int64_t rows,
                                  int cols4) {
    const int64_t total = rows * cols4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / cols4;
        const int c = (int)(i % cols4);
        float4 v = *reinterpret_cast<const float4*>(x + r * ldx + 4 * c);
        v.x = fminf(fmaxf(v.x, -65504.f), 65504.f);
        v.y = fminf(fmaxf(v.y, -65504.f), 65504.f);
        v.z = fminf(fmaxf(v.z, -65504.f), 65504.f);
        v.w = fminf(fmaxf(v.w, -65504.f), 65504.f);
        __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
        uint2 u;
        u.x = *reinterpret_cast<uint32_t*>(&a);
        u.y = *reinterpret_cast<uint32_t*>(&b);
        reinterpret_cast<uint2*>(y)[i] = u;
    }
}

// fp32 -> split fp16 operand for the 3-product GEMM: x = hi + lo with hi = fp16(x), lo = fp16(x - hi).
//   activations (weights = 0): row [hi | hi | lo]      weights (weights = 1): row [hi | lo | hi]
// so that A'.W'^T = hi.hi + hi.lo + lo.hi: every product of two fp16 values is exact in the fp32 accumulator and the
// dropped lo.lo term is 2^-22 relative — fp32-class accuracy on the fp16 tensor pipe.
__global__ void split3_f16_kernel(const float* __restrict__ x, int64_t ldx, __half* __restrict__ y, int64_t rows, int K4,
                                  int weights) {
    const int64_t total = rows * K4;
    const int K = K4 * 4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / K4;
        const int c = (int)(i % K4) * 4;
        const float4 v = *reinterpret_cast<const float4*>(x + r * ldx + c);
        const float f[4] = {v.x, v.y, v.z, v.w};
        __half hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float s = fminf(fmaxf(f[e], -65504.f), 65504.f);
            hi[e] = __float2half_rn(s);
            lo[e] = __float2half_rn(s - __half2float(hi[e]));
        }
        __half* row = y + r * (int64_t)(3 * K);
        const uint2 H = *reinterpret_cast<const uint2*>(hi), L = *reinterpret_cast<const uint2*>(lo);
        *reinterpret_cast<uint2*>(row + c) = H;
        *reinterpret_cast<uint2*>(row + K + c) = weights ? L : H;
        *reinterpret_cast<uint2*>(row + 2 * K + c) = weights ? H : L;
    }
}

// ------------------------------------------------------------------------------------- host helpers
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 2-D tensor [rows, cols] with row pitch ld (elements); box = [box_cols, box_rows] with a 128-byte inner extent,
// 128-byte swizzle
int make_map(CUtensorMap* map, const void* base, bool f32, int64_t rows, int64_t cols, int64_t ld, int box_cols,
             int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return CONE_ERR_CUDA;
    }
    const int esz = f32 ? 4 : 2;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * esz};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                    const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) for [%lld x %lld] ld %lld", (int)r, (long long)rows,
                  (long long)cols, (long long)ld);
        return CONE_ERR_CUDA;
    }
    return CONE_OK;
}

template <int BN, bool WRES, int MODE>
constexpr size_t tc_smem_bytes() {
    constexpr size_t tail = 64 * sizeof(uint64_t) + 8 * 32 * 8 + 64;  // barriers, LayerNorm partials, TMEM slot
    constexpr size_t b_bytes = (size_t)BN * BK * 2;
    if (WRES) return (size_t)WRES_STAGES * A_BYTES + WRES_KBLOCKS * b_bytes + 8 * BOX_BYTES + tail;
    if (MODE == EPI_PLAIN16) return (size_t)FAST_STAGES * (A_BYTES + b_bytes) + 8 * BOX_BYTES + tail;
    return 1024 + (size_t)STAGES * (A_BYTES + b_bytes) + 16 * BOX_BYTES + tail;
}
static_assert(tc_smem_bytes<256, true, EPI_GENERIC>() <= 232448, "weights-resident layout exceeds 227 KB");
static_assert(tc_smem_bytes<256, false, EPI_PLAIN16>() <= 232448, "4-stage ring layout exceeds 227 KB");

}  // namespace

struct TcWeights {
    struct W16 {
        __half* ptr;
        CUtensorMap map;
        int N, K, BN;  // K = columns of the fp16 copy (3 x the fp32 K for a split copy)
    };
    // fp16 copies of nn.Linear weights, keyed by the fp32 device pointer and the kind of copy (plain / split)
    std::map<std::pair<const float*, int>, W16> cache;
    char* scratch = nullptr;
    size_t scratch_bytes = 0;
    int num_sms = kNumSMs;
};

int tc_weights_create(TcWeights** out, cudaStream_t) {
    TcWeights* t = new TcWeights();
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&t->num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (!encode_fn()) {
        delete t;
        set_error("CONE_PREC_TC needs cuTensorMapEncodeTiled (driver too old?)");
        return CONE_ERR_CUDA;
    }
    *out = t;
    return CONE_OK;
}

void tc_weights_destroy(TcWeights* t) {
    if (!t) return;
    for (auto& kv : t->cache) cudaFree(kv.second.ptr);
    delete t;
}

size_t tc_scratch_bytes(int64_t max_rows, int max_k) { return (size_t)max_rows * max_k * 2 + 256; }

void tc_set_scratch(TcWeights* t, void* scratch, size_t bytes) {
    if (!t) return;
    t->scratch = (char*)scratch;
    t->scratch_bytes = bytes;
}

bool tc_gemm_supported(int64_t M, int N, int K) { return M >= 1 && (K % BK) == 0 && (N % 128) == 0; }

int f32_to_f16(const float* x, int64_t ldx, __half* y, int64_t rows, int cols, cudaStream_t s) {
    if (rows == 0) return CONE_OK;
    const int64_t total = rows * (cols / 4);
    const int64_t want = cdiv64(total, 256);
    const unsigned grid = (unsigned)(want < (int64_t)kNumSMs * 16 ? want : (int64_t)kNumSMs * 16);
    ProfScope ps(s, P_CONVERT, 0.0, 6.0 * (double)rows * cols);
    f32_to_f16_kernel<<<grid, 256, 0, s>>>(x, ldx, y, rows, cols / 4);
    CONE_LAUNCH_CHECK("f32_to_f16");
    return CONE_OK;
}

int f32_to_f16_rows(const float* x, int64_t ldx, uint16_t* y, int64_t rows, int cols, cudaStream_t s) {
    return f32_to_f16(x, ldx, reinterpret_cast<__half*>(y), rows, cols, s);
}

static int split3(const float* x, int64_t ldx, __half* y, int64_t rows, int K, int weights, cudaStream_t s) {
    if (rows == 0) return CONE_OK;
    const int64_t total = rows * (K / 4);
    const int64_t want = cdiv64(total, 256);
    const unsigned grid = (unsigned)(want < (int64_t)kNumSMs * 16 ? want : (int64_t)kNumSMs * 16);
    ProfScope ps(s, P_CONVERT, 0.0, 10.0 * (double)rows * K);
    split3_f16_kernel<<<grid, 256, 0, s>>>(x, ldx, y, rows, K / 4, weights);
    CONE_LAUNCH_CHECK("split3_f16");
    return CONE_OK;
}

int split3_f16_rows(const float* x, int64_t ldx, uint16_t* y, int64_t rows, int K, cudaStream_t s) {
    CONE_REQUIRE((K & 3) == 0 && (ldx & 3) == 0, "split3_f16_rows: K and ldx must be multiples of 4");
    return split3(x, ldx, reinterpret_cast<__half*>(y), rows, K, 0, s);
}

// K = the fp32 weight's K; split = 1 makes the [N, 3 K] split copy of the 3-product GEMM
static int get_w16(TcWeights* t, const float* W, int N, int K, int split, cudaStream_t s, const TcWeights::W16** out) {
    const auto key = std::make_pair(W, split);
    const int K16 = split ? 3 * K : K;
    auto it = t->cache.find(key);
    if (it == t->cache.end() || it->second.N != N || it->second.K != K16) {
        TcWeights::W16 w{};
        w.N = N;
        w.K = K16;
        w.BN = (N % 256 == 0) ? 256 : 128;
        CONE_CUDA(cudaMalloc(&w.ptr, (size_t)N * K16 * 2));
        if (split) CONE_TRY(split3(W, K, w.ptr, N, K, 1, s));
        else CONE_TRY(f32_to_f16(W, K, w.ptr, N, K, s));
        CONE_TRY(make_map(&w.map, w.ptr, false, N, K16, K16, BK, w.BN));
        if (it != t->cache.end()) cudaFree(it->second.ptr);
        t->cache[key] = w;
        it = t->cache.find(key);
    }
    *out = &it->second;
    return CONE_OK;
}

// The fp32 master weights changed in place (cone_weights_update): re-round every cached fp16 copy into its existing
// buffer, so device pointers and TMA descriptors (and any CUDA graph that captured them) stay valid.
int tc_weights_refresh(TcWeights* t, cudaStream_t s) {
    if (!t) return CONE_OK;
    for (auto& kv : t->cache) {
        if (kv.first.second) CONE_TRY(split3(kv.first.first, kv.second.K / 3, kv.second.ptr, kv.second.N, kv.second.K / 3, 1, s));
        else CONE_TRY(f32_to_f16(kv.first.first, kv.second.K, kv.second.ptr, kv.second.N, kv.second.K, s));
    }
    return CONE_OK;
}

int tc_weight_f16(TcWeights* t, const float* W, int N, int K, cudaStream_t s, const uint16_t** out) {
    CONE_REQUIRE(t != nullptr, "tc_weight_f16: tensor-core weights not initialised");
    const TcWeights::W16* w = nullptr;
    CONE_TRY(get_w16(t, W, N, K, 0, s, &w));
    *out = reinterpret_cast<const uint16_t*>(w->ptr);
    return CONE_OK;
}

int tc_make_map(CUtensorMap* map, const void* base, bool f32, int64_t rows, int64_t cols, int64_t ld, int box_cols,
                int box_rows) {
    return make_map(map, base, f32, rows, cols, ld, box_cols, box_rows);
}

int tc_num_sms(const TcWeights* t) { return t ? t->num_sms : kNumSMs; }

int tc_gemm_run(TcWeights* t, const TcGemmArgs& g, cudaStream_t s) {
    CONE_REQUIRE(t != nullptr, "tc_gemm: tensor-core weights not initialised");
    CONE_REQUIRE(!(g.split3 && g.wsplit), "tc_gemm: split3 and wsplit are exclusive");
    const int K16 = g.split3 ? 3 * g.K : (g.wsplit ? 2 * g.K : g.K);  // contraction length on the tensor pipe
    CONE_REQUIRE(tc_gemm_supported(g.M, g.N, K16), "tc_gemm: unsupported shape M=%lld N=%d K=%d", (long long)g.M, g.N, K16);
    CONE_REQUIRE(g.M < (int64_t)1 << 31, "tc_gemm: more than 2^31 rows");
    CONE_REQUIRE((g.lda % 8) == 0 && (reinterpret_cast<uintptr_t>(g.A16) & 15) == 0, "tc_gemm: A must be 16-byte aligned");
    CONE_REQUIRE(g.C32 == nullptr || ((g.ldc32 % 4) == 0 && (reinterpret_cast<uintptr_t>(g.C32) & 15) == 0), "tc_gemm: C32 alignment");
    CONE_REQUIRE(g.C16 == nullptr || ((g.ldc16 % 8) == 0 && (reinterpret_cast<uintptr_t>(g.C16) & 15) == 0), "tc_gemm: C16 alignment");
    CONE_REQUIRE(g.R16 == nullptr || ((g.ldr16 % 8) == 0 && (reinterpret_cast<uintptr_t>(g.R16) & 15) == 0), "tc_gemm: R16 alignment");
    CONE_REQUIRE(g.R32 == nullptr || ((g.ldr32 % 4) == 0 && (reinterpret_cast<uintptr_t>(g.R32) & 15) == 0), "tc_gemm: R32 alignment");
    const TcWeights::W16* w = nullptr;
    CONE_TRY(get_w16(t, g.W, g.N, g.K, g.split3 || g.wsplit, s, &w));
    CONE_REQUIRE(g.ln_g == nullptr || g.N == w->BN, "tc_gemm: fused LayerNorm needs the whole row in one tile (N=%d)", g.N);
    CONE_REQUIRE(!(g.C16 && g.C32 && g.R16 && g.ln_g == nullptr),
                 "tc_gemm: fp16 + fp32 outputs with an fp16 residual need the LayerNorm epilogue");
    CUtensorMap mapA, mapR, mapC16, mapC32;
    CONE_TRY(make_map(&mapA, g.A16, false, g.M, g.wsplit ? g.K : K16, g.lda, BK, BM));
    mapR = mapA;
    mapC16 = mapA;
    mapC32 = mapA;  // placeholders when unused (never dereferenced)
    if (g.R16) CONE_TRY(make_map(&mapR, g.R16, false, g.M, g.N, g.ldr16, 64, 32));
    if (g.C16) CONE_TRY(make_map(&mapC16, g.C16, false, g.M, g.N, g.ldc16, 64, 32));
    if (g.C32) CONE_TRY(make_map(&mapC32, g.C32, true, g.M, g.N, g.ldc32, 32, 32));
    TcEpilogue ep{};
    ep.bias = g.bias;
    ep.R32 = g.R32;
    ep.ldr32 = g.ldr32;
    ep.has_r16 = g.R16 != nullptr;
    ep.has_c16 = g.C16 != nullptr;
    ep.has_c32 = g.C32 != nullptr;
    ep.relu = g.relu;
    ep.ln_g = g.ln_g;
    ep.ln_b = g.ln_b;
    ep.ln_eps = 1e-5f;
    const int64_t m_tiles = cdiv64(g.M, BM);
    const int n_tiles = g.N / w->BN;
    const int64_t tiles = m_tiles * n_tiles;
    // weights-resident variant: K = 256, BN = 256, at most num_sms / n_tiles CTAs per n-tile
    const bool wres = w->BN == 256 && K16 == WRES_KBLOCKS * BK && n_tiles <= t->num_sms && !(g.C16 && g.C32) && !g.wsplit;
    const int a_wrap = g.wsplit ? g.K : K16;
    unsigned grid = (unsigned)(tiles < t->num_sms ? tiles : t->num_sms);
    if (wres) {
        const int64_t per_n = t->num_sms / n_tiles;
        grid = (unsigned)((m_tiles < per_n ? m_tiles : per_n) * n_tiles);
    }
    const double mn = (double)g.M * g.N;
    // FLOPs are the ALGORITHMIC count (2 M N K of the fp32 problem): the split GEMMs execute 2-3x that on the tensor pipe
    ProfScope ps(s, P_GEMM_TC, 2.0 * mn * g.K,
                 2.0 * ((double)g.M * K16 + (double)g.N * K16) + (g.C32 ? 4.0 : 0.0) * mn + (g.C16 ? 2.0 : 0.0) * mn +
                     (g.R32 ? 4.0 : 0.0) * mn + (g.R16 ? 2.0 : 0.0) * mn);
#define CONE_TC_LAUNCH(BNV, WR, MD)                                                                                 \
    do {                                                                                                                \
        static bool attr = false;                                                                                       \
        if (!attr) {                                                                                                    \
            CONE_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<BNV, WR, MD>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                           (int)tc_smem_bytes<BNV, WR, MD>()));                                         \
            attr = true;                                                                                                \
        }                                                                                                               \
        tc_gemm_kernel<BNV, WR, MD><<<grid, TC_THREADS, tc_smem_bytes<BNV, WR, MD>(), s>>>(                             \
            mapA, w->map, mapR, mapC16, mapC32, ep, g.M, g.N, K16, a_wrap);                                                     \
    } while (0)
    const bool plain16 = w->BN == 256 && g.C16 && !g.C32 && !g.R16 && !g.R32 && !g.ln_g && g.bias;
    const bool ln16 = w->BN == 256 && g.C16 && !g.C32 && g.R16 && !g.R32 && g.ln_g && g.bias && !g.relu;
    if (w->BN != 256) {
        CONE_TC_LAUNCH(128, false, EPI_GENERIC);
    } else if (wres) {
        if (plain16) CONE_TC_LAUNCH(256, true, EPI_PLAIN16);
        else if (ln16) CONE_TC_LAUNCH(256, true, EPI_LN16);
        else CONE_TC_LAUNCH(256, true, EPI_GENERIC);
    } else {
        if (plain16) CONE_TC_LAUNCH(256, false, EPI_PLAIN16);
        else if (ln16) CONE_TC_LAUNCH(256, false, EPI_LN16);
        else CONE_TC_LAUNCH(256, false, EPI_GENERIC);
    }
#undef CONE_TC_LAUNCH
    CONE_LAUNCH_CHECK("tc_gemm");
    return CONE_OK;
}

int tc_gemm(TcWeights* t, const float* x, int64_t ldx, int64_t M, const float* W, const float* b, int N, int K, float* y,
            int64_t ldy, int relu, const float* R, int64_t ldr, cudaStream_t s) {
    CONE_REQUIRE(t != nullptr, "tc_gemm: tensor-core weights not initialised");
    const size_t need = (size_t)M * K * 2;
    CONE_REQUIRE(t->scratch != nullptr && t->scratch_bytes >= need,
                 "tc_gemm: operand staging needs %zu bytes of workspace, %zu available", need, t->scratch_bytes);
    __half* a16 = reinterpret_cast<__half*>(t->scratch);
    CONE_TRY(f32_to_f16(x, ldx, a16, M, K, s));
    TcGemmArgs g;
    g.A16 = reinterpret_cast<const uint16_t*>(a16);
    g.lda = K;
    g.M = M;
    g.W = W;
    g.bias = b;
    g.N = N;
    g.K = K;
    g.C32 = y;
    g.ldc32 = ldy;
    g.relu = relu;
    g.R32 = R;
    g.ldr32 = ldr;
    return tc_gemm_run(t, g, s);
}

}  // namespace cone

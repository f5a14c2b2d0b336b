// Encoder self-attention on tcgen05 (CONE_PREC_TC): nn.MultiheadAttention of `TransformerEncoderLayer.forward_post`
// (cone/transformer.py:239), 8 heads of 32 channels over one window of S = Lv + Lt <= 190 rows, key-padding mask.
//
// Why a second attention kernel: the mma.sync version (attention.cu, one CTA per (window, head)) spends 4 700 SM-cycles per
// (window, head) — LSU-bound (l1tex 86 %: every q / k / v / position value goes global -> register -> shared -> ldmatrix,
// the score block lives in mma.sync fragments that need quad shuffles for every row maximum) — against a floor of ~1 600
// cycles set by the 22 500 exponentials per head.  Here:
//   * q, k and v arrive by TMA, no register staging: q and k as boxes of a head PAIR (64 channels = 128-byte rows, 128-byte
//     swizzle; the two heads are the K sub-blocks of one swizzle atom), v as 64-byte-swizzled boxes of one head; narrow rows
//     are what limits TMA here (ncu: with five 64-byte-row boxes per head the kernel ran at the TMA unit's ~8 cycles per row);
//     two helper warps add the position-projection rows (read from the L2-resident table) to q and k in place;
//   * S = Q.K^T runs on tcgen05 (M = 128 per tile, N = padded key count, K = 32) into TMEM; softmax is ONE THREAD PER ROW
//     on `tcgen05.ld` data: no shuffles, no fragments; P goes back to shared memory as the fp16 A operand (K-major,
//     128-byte swizzle) and O = P.V is a second tcgen05 product with V used as it lands (MN-major B operand, no transpose);
//   * the heads of a window are pipelined: TMA of head h+1, QK^T of head h+1 and P.V of head h overlap the softmax of head h.
// One persistent CTA per SM, 24 warps: 0 TMA producer, 1 MMA issuer (whole warp, elect.sync), 2-3 position add, 4-23 softmax /
// output.  The softmax is exponential- and latency-bound (22 500 ex2 per head against ~1 000 tensor-pipe cycles), so the score
// rows are split by KEY RANGE over four warps per TMEM lane quarter (a warp reads only the lanes of its quarter = its SM
// sub-partition): warps 4-19 take rows 0-127 (quarter = warp & 3, key range g = (warp - 4) / 4); rows 128-159 — a fifth
// 32-row group that would otherwise double the load of one sub-partition — are computed by FOUR score MMAs, one per key
// range, each with its A operand shifted by 32 g rows so that range g of these rows lands in lane quarter g: warps 20-23 take
// one range each, and every sub-partition runs five warps with the same work.  Partial row maxima / sums go through shared
// memory (named barrier per group of four warps).
#include <cuda.h>
#include <cuda_fp16.h>
#include <math_constants.h>
#include <stdlib.h>

#include "kernels.h"
#include "tc_gemm.h"
#include "tc_ptx.cuh"

namespace cone {

using namespace ptx;

namespace {

constexpr int AT_THREADS = 768;
constexpr int AT_G = 4;         // key ranges = softmax warps per row group
constexpr int AT_HD = 32;       // head dim
constexpr int AT_VROWB = 64;    // bytes per v row of one head (32 fp16): the 64-byte swizzle span
constexpr int AT_QROWB = 128;   // bytes per q / k row of a head pair: the 128-byte swizzle span
constexpr int AT_MAX_NKP = 192; // padded key count supported (TMEM: 2 x NKP + 128 <= 512)
constexpr int AT_NQS = 2;       // q|k stages (head pairs in flight)
constexpr int AT_NVS = 3;       // v stages (heads in flight)

// descriptor hi word for 64-byte-swizzled operands: SBO = 8 rows x 64 B = 512 B, version 1, layout SWIZZLE_64B (= 4)
constexpr uint32_t kDescHiSw64 = (uint32_t)(512 >> 4) | (1u << 14) | (4u << 29);

__device__ __forceinline__ void umma_f16_desc(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                              uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
        "}" ::"r"(tmem_d),
        "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct AtParams {
    __half* out;             // [B S, ldo]
    int64_t ldo;
    const int32_t* vlen;
    const int32_t* tlen;
    const int64_t* vid_base; // layer 0: first frame row of each window (null = dense rows)
    const int64_t* txt_base;
    const __half* pos;       // position projection [(table_lv + 1) table_lv, 512] fp16: pos.Wq^T | pos.Wk^T
    int64_t B;
    int Lv, Lt, table_lv;
    int nkp;                 // padded key count (multiple of 16)
    int ntile;               // 1 or 2 query tiles of 128 rows
    int indirect;
    int tpad;                // layer 0 with an odd Lv: the token rows start one row later (v rows are 64 bytes, TMA destinations 128-byte aligned)
};

// shared-memory plan (bytes), all multiples of 1 KB
struct AtPlan {
    int qbuf;    // q or k of a head pair: nkp x 128
    int qstage;  // q | k  (rows 128.. of q as an M = 128 operand run on into k: finite values, unused result rows)
    int vbuf;    // v of one head: nkp x 64
    // One P set = [P of rows 128..: two k-blocks of [32 x 128 B] + a 64-byte-swizzled tail block [32 x 64 B] for keys 128-159]
    //             [P of rows 0-127: two k-blocks of [128 x 128 B] + tail block [128 x 64 B]].
    // The M = 128 operand of rows 128.. reads on into the blocks behind it (finite values, unused result rows).
    int p1_tail, p0, p0_tail, pset;
    int off_v, off_p, off_bias, off_x, off_bar, total;
};
// key range g of a window with `units` 16-key units: first unit and unit count
__host__ __device__ inline void at_range(int units, int g, int* u0, int* un) {
    const int ub = units / AT_G, ur = units % AT_G;
    *u0 = g * ub + (g < ur ? g : ur);
    *un = ub + (g < ur ? 1 : 0);
}
__host__ __device__ inline AtPlan at_plan(int nkp, int ntile) {
    AtPlan p;
    p.qbuf = ((nkp * AT_QROWB + 1023) / 1024) * 1024;
    p.qstage = 2 * p.qbuf;
    p.vbuf = ((nkp * AT_VROWB + 1023) / 1024) * 1024;
    const int nblk = nkp > 64 ? 2 : 1;       // 128-byte-swizzled k-blocks of 64 keys (keys 128.. go to the tail block)
    p.p1_tail = ntile == 2 ? nblk * 4096 : 0;
    p.p0 = ntile == 2 ? p.p1_tail + 2048 : 0;
    p.p0_tail = p.p0 + nblk * 16384;
    p.pset = p.p0_tail + (nkp > 128 ? 8192 : 0);
    p.off_v = AT_NQS * p.qstage;
    p.off_p = p.off_v + AT_NVS * p.vbuf;      // two P sets: the softmax of head h + 1 writes while P.V of head h reads
    p.off_bias = p.off_p + 2 * p.pset;        // [20][64] fp32 additive key mask of the warp's key range (one copy per warp)
    p.off_x = p.off_bias + 20 * 256;          // partial maxima [AT_G][160] and partial sums [2][AT_G][160] (row groups 0-3, 4)
    p.off_bar = p.off_x + 3 * AT_G * 160 * 4;
    p.total = p.off_bar + 256;
    return p;
}

__global__ void __launch_bounds__(AT_THREADS, 1)
enc_attn_tc_kernel(const __grid_constant__ CUtensorMap tmQK,    // head-pair boxes of dense rows [B S, 768] (layer 1) or frames [n_frames, 768]
                   const __grid_constant__ CUtensorMap tmQKTok, // ... of tokens [n_tok, 768] (layer 0 only)
                   const __grid_constant__ CUtensorMap tmV,     // one-head boxes of the same tensors
                   const __grid_constant__ CUtensorMap tmVTok,
                   AtParams P) {
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    const AtPlan pl = at_plan(P.nkp, P.ntile);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + pl.off_bar);
    uint64_t* qk_full = bars;                       // [AT_NQS] q and k of a head pair have landed
    uint64_t* qk_ready = bars + AT_NQS;             // [AT_NQS] position rows added to q and k
    uint64_t* qk_free = bars + 2 * AT_NQS;          // [AT_NQS] Q.K^T of the pair's second head has finished reading q and k
    uint64_t* v_full = bars + 3 * AT_NQS;           // [AT_NVS] v of a head has landed
    uint64_t* v_free = v_full + AT_NVS;             // [AT_NVS] P.V has finished reading v
    uint64_t* s_full = v_free + AT_NVS;             // S = Q.K^T complete
    uint64_t* p_ready = s_full + 1;                 // P in shared memory, S read out of TMEM
    uint64_t* p_free = s_full + 2;                  // [2] P.V has finished reading the P set
    uint64_t* o_full = s_full + 4;                  // [2]
    uint64_t* o_free = s_full + 6;                  // [2]
    uint64_t* s_free = s_full + 8;                  // S copied to registers: the next head's Q.K^T may overwrite it
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 9);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = P.Lv + P.Lt, nkp = P.nkp, NT = P.ntile;
    const int n_soft = (NT == 2) ? 20 : 16;  // softmax warps in use
    const int n_out = (NT == 2) ? 17 : 16;   // ... of which read O (rows 128.. are written out by one warp)

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmQK)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmV)) : "memory");
        for (int i = 0; i < AT_NQS; ++i) {
            mbar_init(&qk_full[i], 1);
            mbar_init(&qk_ready[i], 2);
            mbar_init(&qk_free[i], 1);
        }
        for (int i = 0; i < AT_NVS; ++i) {
            mbar_init(&v_full[i], 1);
            mbar_init(&v_free[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&o_full[i], 1);
            mbar_init(&o_free[i], n_out);
            mbar_init(&p_free[i], 1);
        }
        mbar_init(s_full, 1);
        mbar_init(p_ready, n_soft);
        mbar_init(s_free, n_soft);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // Rows that TMA never writes must hold finite values: v rows >= S multiply P = 0 (0 x NaN would poison the output)
    for (int i = threadIdx.x; i < pl.off_p / 16; i += AT_THREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (*tmem_slot != 0u) __trap();  // the whole TMEM belongs to this CTA: addresses below are literals
    const uint32_t tmemS = 0u;                          // [NT][nkp] fp32 scores
    const uint32_t tmemO = (uint32_t)(NT * nkp);        // [2][NT][32] fp32 outputs, double buffered across heads

    const int64_t w_begin = blockIdx.x, w_step = gridDim.x;

    // 768 threads leave 80 registers each; the softmax warps hold their 48 scores in registers across the exchange of the row
    // maxima, the four service warps (one warpgroup) need few: 128 x 40 registers move to the five softmax warpgroups
    // (each setmaxnreg sits inside its role's branch: code reachable from both would be compiled for the smaller budget)
    if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 0) {
        if (lane == 0) {  // ------------------------------------------------------------------------------ TMA producer
            int qs = 0, vs = 0;
            uint32_t qph = 0, vph = 0;
            const uint32_t bytes_q = (uint32_t)(S * AT_QROWB), bytes_v = (uint32_t)(S * AT_VROWB);
            for (int64_t w = w_begin; w < P.B; w += w_step) {
                const int64_t vb = P.indirect ? P.vid_base[w] : w * S;
                const int64_t tb = P.indirect ? P.txt_base[w] : 0;
                for (int h = 0; h < 8; ++h) {
                    if ((h & 1) == 0) {  // q and k of heads h, h + 1: columns m * 256 + h * 32 .. + 64
                        uint8_t* base = smem + qs * pl.qstage;
                        mbar_wait(&qk_free[qs], qph ^ 1);
                        mbar_expect_tx(&qk_full[qs], 2 * bytes_q);
                        for (int m = 0; m < 2; ++m) {
                            tma_load_2d(base + m * pl.qbuf, &tmQK, &qk_full[qs], m * 256 + h * AT_HD, (int)vb);
                            if (P.indirect)
                                tma_load_2d(base + m * pl.qbuf + (P.Lv + P.tpad) * AT_QROWB, &tmQKTok, &qk_full[qs], m * 256 + h * AT_HD, (int)tb);
                        }
                        if (++qs == AT_NQS) { qs = 0; qph ^= 1; }
                    }
                    uint8_t* vbase = smem + pl.off_v + vs * pl.vbuf;
                    mbar_wait(&v_free[vs], vph ^ 1);
                    mbar_expect_tx(&v_full[vs], bytes_v);
                    tma_load_2d(vbase, &tmV, &v_full[vs], 512 + h * AT_HD, (int)vb);
                    if (P.indirect) tma_load_2d(vbase + (P.Lv + P.tpad) * AT_VROWB, &tmVTok, &v_full[vs], 512 + h * AT_HD, (int)tb);
                    if (++vs == AT_NVS) { vs = 0; vph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {  // --------------------------------------------------------------------------------- MMA issuer
        const uint32_t idS = (1u << 4) | ((uint32_t)(nkp >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);                 // K-major A, B
        const uint32_t idO = (1u << 4) | (1u << 16) | ((uint32_t)(AT_HD >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);  // B (= V) MN-major
        const uint32_t base_lo = desc_lo_sw128(smem_u32(smem));  // (address >> 4) | LBO = 1
        uint32_t it = 0;
        int qs = 0, vs = 0, vs_prev = 0;
        uint32_t qph = 0, vph = 0, vph_prev = 0;
        auto issue_pv = [&](uint32_t itp, int vsp, uint32_t vphp, bool wait_p) {  // O = P . V of the head issued at iteration itp
            const int ob = itp & 1;
            if (wait_p) mbar_wait(p_ready, itp & 1);  // (the main loop has already waited: never wait twice on a parity)
            mbar_wait(&v_full[vsp], vphp);
            mbar_wait(&o_free[ob], ((itp >> 1) & 1) ^ 1);
            tc_fence_after();
            if (elect_one_sync()) {
                const uint32_t v_lo = base_lo + (uint32_t)((pl.off_v + vsp * pl.vbuf) >> 4);
                const uint32_t set_lo = base_lo + (uint32_t)((pl.off_p + ob * pl.pset) >> 4);  // P set of this head
                for (int t = 0; t < NT; ++t) {
                    const uint32_t d = tmemO + (uint32_t)((ob * NT + t) * AT_HD);
                    const bool t1 = NT == 2 && t == 1;  // rows 128..
                    const uint32_t blk_lo = set_lo + (uint32_t)((t1 ? 0 : pl.p0) >> 4), blk_sz = t1 ? 4096u : 16384u;
                    const uint32_t tail_lo = set_lo + (uint32_t)((t1 ? pl.p1_tail : pl.p0_tail) >> 4);
                    for (int j = 0; j < nkp / 16; ++j) {  // 16 keys per MMA: A = P[:, 16 j ..], B = V[16 j .., :]
                        const uint32_t b = v_lo + (uint32_t)((j * 16 * AT_VROWB) >> 4);
                        if (j < 8) umma_f16_desc(d, blk_lo + (((uint32_t)(j >> 2) * blk_sz + (uint32_t)(j & 3) * 32u) >> 4), kDescHiSw128, b, kDescHiSw64, idO, j > 0 ? 1u : 0u);
                        else umma_f16_desc(d, tail_lo + (uint32_t)(((j - 8) * 32) >> 4), kDescHiSw64, b, kDescHiSw64, idO, 1u);
                    }
                }
                umma_commit(&o_full[ob]);
                umma_commit(&p_free[ob]);
                umma_commit(&v_free[vsp]);
            }
            __syncwarp();
        };
        for (int64_t w = w_begin; w < P.B; w += w_step) {
            for (int h = 0; h < 8; ++h, ++it) {
                if ((h & 1) == 0) mbar_wait(&qk_ready[qs], qph);
                if (it > 0) mbar_wait(s_free, (it - 1) & 1);  // S of the previous head is in the softmax warps' registers
                tc_fence_after();
                if (elect_one_sync()) {
                    // the head is the (h & 1)-th 64-byte half of the pair's 128-byte rows: K sub-blocks 2 (h & 1), 2 (h & 1) + 1
                    const uint32_t q_lo = base_lo + (uint32_t)((qs * pl.qstage) >> 4) + 4 * (h & 1), k_lo = q_lo + (uint32_t)(pl.qbuf >> 4);
                    for (int k = 0; k < 2; ++k) umma_f16_desc(tmemS, q_lo + 2 * k, kDescHiSw128, k_lo + 2 * k, kDescHiSw128, idS, k > 0 ? 1u : 0u);
                    if (NT == 2) {
                        // rows 128..: key range g against the Q rows starting at 128 - 32 g -> these rows land in lane quarter g
                        for (int g = 0; g < AT_G; ++g) {
                            int u0, un;
                            at_range(nkp >> 4, g, &u0, &un);
                            if (un == 0) continue;
                            const uint32_t idg = (1u << 4) | ((uint32_t)((un * 16) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
                            const uint32_t a = q_lo + (uint32_t)(((128 - 32 * g) * AT_QROWB) >> 4), b = k_lo + (uint32_t)((u0 * 16 * AT_QROWB) >> 4);
                            for (int k = 0; k < 2; ++k)
                                umma_f16_desc(tmemS + (uint32_t)(nkp + u0 * 16), a + 2 * k, kDescHiSw128, b + 2 * k, kDescHiSw128, idg, k > 0 ? 1u : 0u);
                        }
                    }
                    umma_commit(s_full);
                    if (h & 1) umma_commit(&qk_free[qs]);
                }
                __syncwarp();
                if (h & 1) {
                    if (++qs == AT_NQS) { qs = 0; qph ^= 1; }
                }
                if (it > 0) issue_pv(it - 1, vs_prev, vph_prev, true);  // while the softmax warps reduce this head's scores
                vs_prev = vs;
                vph_prev = vph;
                if (++vs == AT_NVS) { vs = 0; vph ^= 1; }
            }
        }
        if (it > 0) issue_pv(it - 1, vs_prev, vph_prev, true);
    } else {  // ---------------------------------------------------------------------------------- position add (2 warps)
        // q += pos.Wq^T, k += pos.Wk^T over the first Lv rows of both heads of the pair: 16-byte chunks (8 channels), the
        // position rows come from the table (L2-resident: 2 x 64 channels x Lv rows per pair), 4 loads in flight per thread
        const int t = threadIdx.x - 64;  // 0..63
        int qs = 0;
        uint32_t qph = 0;
        const int nchunk = P.Lv * 8;     // per matrix
        for (int64_t w = w_begin; w < P.B; w += w_step) {
            const __half* ptab = P.pos + (int64_t)P.vlen[w] * P.table_lv * 512;
            for (int hp = 0; hp < 4; ++hp) {
                uint8_t* base = smem + qs * pl.qstage;
                bool waited = false;
                for (int i0 = t; i0 < 2 * nchunk; i0 += 64 * 4) {
                    uint4 pv[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int i = i0 + u * 64;
                        if (i < 2 * nchunk) {
                            const int m = i >= nchunk, c = m ? i - nchunk : i;
                            pv[u] = __ldg(reinterpret_cast<const uint4*>(ptab + (int64_t)(c >> 3) * 512 + m * 256 + hp * 64 + (c & 7) * 8));
                        }
                    }
                    if (!waited) {
                        mbar_wait(&qk_full[qs], qph);
                        waited = true;
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int i = i0 + u * 64;
                        if (i < 2 * nchunk) {
                            const int m = i >= nchunk, c = m ? i - nchunk : i;
                            uint4* dst = reinterpret_cast<uint4*>(base + m * pl.qbuf + sw128(c >> 3, c & 7));
                            const uint4 a = *dst;
                            uint4 r;
                            const __half2* ah = reinterpret_cast<const __half2*>(&a);
                            const __half2* bh = reinterpret_cast<const __half2*>(&pv[u]);
                            __half2* rh = reinterpret_cast<__half2*>(&r);
#pragma unroll
                            for (int e = 0; e < 4; ++e) rh[e] = __hadd2(ah[e], bh[e]);  // exact sum, one rounding
                            *dst = r;
                        }
                    }
                }
                if (!waited) mbar_wait(&qk_full[qs], qph);
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&qk_ready[qs]);
                if (++qs == AT_NQS) { qs = 0; qph ^= 1; }
            }
        }
    }
    } else if (warp < 4 + n_soft) {  // ------------------------------------------------------------ softmax / output warps
        asm volatile("setmaxnreg.inc.sync.aligned.u32 88;");
        const int tile = warp >= 20 ? 1 : 0;    // row group: rows 0-127 | rows 128..
        const int quarter = warp & 3;           // TMEM lane quarter (= SM sub-partition) of this warp
        const int g = tile ? quarter : (warp - 4) >> 2;  // key range
        const int xrow = tile ? 128 + lane : quarter * 32 + lane;  // row of this thread in the window tiles (exchange index)
        const int prow_i = tile ? lane : quarter * 32 + lane;      // its row inside the P / O tile of the group
        const int T0 = P.Lv + P.tpad;                              // first token row in the tiles
        // window row (= output row) of the tile row: the video rows, then the token rows; -1 = no row (gap / padding)
        const int row = xrow < P.Lv ? xrow : ((xrow >= T0 && xrow < T0 + P.Lt) ? xrow - P.tpad : -1);
        const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
        int u0, un;
        at_range(nkp >> 4, g, &u0, &un);
        const int k0 = u0 * 16;                                    // first key of this warp's range
        const uint32_t tS = tmemS + lane_base + (uint32_t)(tile * nkp + k0);
        float* bias = reinterpret_cast<float*>(smem + pl.off_bias) + (warp - 4) * 64;  // this warp's copy: no cross-warp hand-off
        float* xmax = reinterpret_cast<float*>(smem + pl.off_x);    // [AT_G][160]
        float* xsum = xmax + AT_G * 160;                            // [2][AT_G][160]
        const int p_blk0 = pl.off_p + (tile ? 0 : pl.p0), p_tail = pl.off_p + (tile ? pl.p1_tail : pl.p0_tail);  // in P set 0
        const int pblk = tile ? 4096 : 16384;                       // bytes per 64-key k-block of the P tile
        const int bar_id = tile ? 5 : 1 + quarter;                  // the four warps that share these rows
        const bool writer = tile == 0 || quarter == 0;              // reads O and writes the output rows
        uint32_t it = 0;
        int64_t w_prev = 0;
        int h_prev = 0;
        auto write_out = [&](uint32_t itp, int64_t wq, int hq) {  // O of head hq of window wq -> global memory
            if (!writer) return;
            const int ob = itp & 1;
            mbar_wait(&o_full[ob], (itp >> 1) & 1);
            tc_fence_after();
            // partial sums of the four key ranges: written before the ranges' p_ready arrivals, which the P.V commit follows
            const float* xs = xsum + ob * AT_G * 160 + xrow;
            const float inv = 1.f / ((xs[0] + xs[160]) + (xs[320] + xs[480]));
            const uint32_t tO = tmemO + lane_base + (uint32_t)((ob * NT + tile) * AT_HD);
            if (tile == 0) {  // 8 of the 32 output columns per warp: the four warps of a quarter write one 64-byte row segment
                float o[8];
                tmem_ld_32x8(tO + 8 * g, o);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&o_free[ob]);
                if (row >= 0)
                    *reinterpret_cast<uint4*>(P.out + (wq * S + row) * P.ldo + hq * AT_HD + 8 * g) =
                        make_uint4(pack_h2(o[0] * inv, o[1] * inv), pack_h2(o[2] * inv, o[3] * inv), pack_h2(o[4] * inv, o[5] * inv),
                                   pack_h2(o[6] * inv, o[7] * inv));
            } else {
                float o[32];
                tmem_ld_32x32(tO, o);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&o_free[ob]);
                if (row >= 0) {
                    uint4* dst = reinterpret_cast<uint4*>(P.out + (wq * S + row) * P.ldo + hq * AT_HD);
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        dst[u] = make_uint4(pack_h2(o[8 * u] * inv, o[8 * u + 1] * inv), pack_h2(o[8 * u + 2] * inv, o[8 * u + 3] * inv),
                                            pack_h2(o[8 * u + 4] * inv, o[8 * u + 5] * inv), pack_h2(o[8 * u + 6] * inv, o[8 * u + 7] * inv));
                }
            }
        };
        for (int64_t w = w_begin; w < P.B; w += w_step) {
            const int vl = P.vlen[w], tl = P.tlen[w];
            const int live_end = T0 + tl;  // keys from here on are padding
            // additive key mask of this warp's range (0 / -inf); only units that hold a masked key read it
            __syncwarp();
            for (int k = lane; k < un * 16; k += 32) {
                const int key = k0 + k;
                bias[k] = (key < vl || (key >= T0 && key < live_end)) ? 0.f : -CUDART_INF_F;
            }
            __syncwarp();
            // 16-key unit at key c: dead = padding only (its P is 0), clean = no masked key (warp-uniform tests)
            auto unit_dead = [&](int c) { return c >= live_end; };
            auto unit_clean = [&](int c) { return (vl >= T0 || c + 16 <= vl || c >= T0) && c + 16 <= live_end; };
            auto add_bias = [&](int u, float* s) {
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    const float4 b = *reinterpret_cast<const float4*>(bias + u * 16 + j);  // broadcast read
                    s[j] += b.x; s[j + 1] += b.y; s[j + 2] += b.z; s[j + 3] += b.w;
                }
            };
            for (int h = 0; h < 8; ++h, ++it) {
                mbar_wait(s_full, it & 1);
                tc_fence_after();
                // this warp's scores (<= 48 keys) move to registers and S is released at once: the next head's Q.K^T runs
                // under this head's softmax
                float sv[2][16];
                // pass 1: maximum of the row over this warp's keys.  Two 16-key units stay in registers; a third unit exists
                // only in key ranges 0 and 1: it is reduced first and read from TMEM again in pass 2, which delays those
                // warps' release of S to the first third of pass 2 (still long before the next head's scores are needed)
                float m0 = -CUDART_INF_F, m1 = -CUDART_INF_F, m2 = -CUDART_INF_F, m3 = -CUDART_INF_F;
                auto max_unit = [&](int u, float* sx) {
                    const int c = k0 + u * 16;
                    if (unit_dead(c)) return;
                    if (!unit_clean(c)) add_bias(u, sx);
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        m0 = fmaxf(m0, sx[j]); m1 = fmaxf(m1, sx[j + 1]); m2 = fmaxf(m2, sx[j + 2]); m3 = fmaxf(m3, sx[j + 3]);
                    }
                };
                const bool third = un > 2 && !unit_dead(k0 + 32);
                if (third) {
                    tmem_ld_32x16(tS + 32, sv[0]);
                    tmem_ld_wait();
                    max_unit(2, sv[0]);
                }
#pragma unroll
                for (int u = 0; u < 2; ++u)
                    if (u < un) tmem_ld_32x16(tS + u * 16, sv[u]);
                tmem_ld_wait();
                if (!third) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(s_free);
                }
#pragma unroll
                for (int u = 0; u < 2; ++u)
                    if (u < un) max_unit(u, sv[u]);
                xmax[g * 160 + xrow] = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
                asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
                const float mx = fmaxf(fmaxf(xmax[xrow], xmax[160 + xrow]), fmaxf(xmax[320 + xrow], xmax[480 + xrow]));
                const float sl2 = 0.17677669529663687f * 1.4426950408889634f;  // 1/sqrt(32) * log2(e)
                const float off = (mx == -CUDART_INF_F) ? 0.f : -mx * sl2;
                // pass 2: p = exp2(s * sl2 - max * sl2) -> fp16 P set it & 1 (K-major; 128-byte swizzle, keys 128.. in the
                // 64-byte-swizzled tail block), fp32 partial row sum
                float sum0 = 0.f, sum1 = 0.f;
                uint8_t* pset = smem + (it & 1) * pl.pset;
                auto exp_unit = [&](int u, float* sx) {
                    const int c = k0 + u * 16;
                    const bool dead = unit_dead(c);
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) {  // 8 keys = one 16-byte chunk of the P row
                        uint4 v = make_uint4(0u, 0u, 0u, 0u);
                        if (!dead) {
                            float* e = sx + hf * 8;
#pragma unroll
                            for (int j = 0; j < 8; ++j) e[j] = ex2_approx(fmaf(e[j], sl2, off));
                            if (hf == 0) sum0 += ((e[0] + e[1]) + (e[2] + e[3])) + ((e[4] + e[5]) + (e[6] + e[7]));
                            else sum1 += ((e[0] + e[1]) + (e[2] + e[3])) + ((e[4] + e[5]) + (e[6] + e[7]));
                            v = make_uint4(pack_h2(e[0], e[1]), pack_h2(e[2], e[3]), pack_h2(e[4], e[5]), pack_h2(e[6], e[7]));
                        }
                        uint8_t* dst;
                        if (c < 128) {
                            dst = pset + p_blk0 + (c >> 6) * pblk + sw128(prow_i, ((c & 63) >> 3) + hf);
                        } else {  // 64-byte rows: 16-byte chunk index XOR (row / 2) % 4
                            const int ch = ((c - 128) >> 3) + hf;
                            dst = pset + p_tail + prow_i * 64 + ((ch ^ ((prow_i >> 1) & 3)) << 4);
                        }
                        *reinterpret_cast<uint4*>(dst) = v;
                    }
                };
                if (it > 1) mbar_wait(&p_free[it & 1], ((it >> 1) & 1) ^ 1);  // P.V of head it - 2 has finished reading this set
                if (un > 0) exp_unit(0, sv[0]);
                if (third) {
                    tmem_ld_32x16(tS + 32, sv[0]);
                    tmem_ld_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(s_free);
                    if (!unit_clean(k0 + 32)) add_bias(2, sv[0]);
                }
                if (un > 1) exp_unit(1, sv[1]);
                if (un > 2) exp_unit(2, sv[0]);  // (a dead third unit only writes its zeros)
                xsum[((it & 1) * AT_G + g) * 160 + xrow] = sum0 + sum1;
                tc_fence_before();
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(p_ready);
                // the previous head's output, while the tensor pipe works on this head's P.V and the next head's scores
                if (it > 0) write_out(it - 1, w_prev, h_prev);
                w_prev = w;
                h_prev = h;
            }
        }
        if (it > 0) write_out(it - 1, w_prev, h_prev);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(0u), "r"(512u) : "memory");
    }
}

// 2-D fp16 map with an inner box of `box_cols` channels: 32 -> 64-byte rows and swizzle, 64 -> 128-byte rows and swizzle
int make_map_box(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows, int box_cols) {
    typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    if (!fn) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return CONE_ERR_CUDA;
    }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, box_cols == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) for [%lld x %lld] ld %lld box %d x %d", (int)r, (long long)rows, (long long)cols,
                  (long long)ld, box_rows, box_cols);
        return CONE_ERR_CUDA;
    }
    return CONE_OK;
}

}  // namespace

bool enc_attn_tc_supported(int Lv, int Lt, int d_model, int nheads) {
    const int S = Lv + Lt + (Lv & 1), nkp = (S + 15) & ~15, nt = (S + 127) / 128;
    static int env = -1;
    if (env < 0) {
        // opt-in: on the benchmark windows (150 rows, 8 heads of 32) this kernel runs at 3.1 ms per launch against 2.75 ms of
        // the mma.sync kernel (profiles/r02_notes.md §5: the exponentials, not the products, bound both)
        const char* e = getenv("CONE_ATTN_TC");
        env = (e && e[0] == '1') ? 1 : 0;
    }
    return env == 1 && d_model == 256 && nheads == 8 && (nkp <= 128 || nkp == 160) && nt <= 2 && Lv <= 256 && Lt >= 1 && Lt <= 256 &&
           nt * nkp + 2 * nt * AT_HD <= 512 && at_plan(nkp, nt).total <= 232448;
}

int enc_attn_tc_run(const void* qkv, int64_t rows, void* o, int64_t ldo, const int32_t* vlen, const int32_t* tlen, int64_t B, int Lv,
                    int Lt, const void* posqk16, int table_lv, const void* token_qkv, int64_t n_tok, const int64_t* vid_base,
                    const int64_t* txt_base, int num_sms, cudaStream_t s) {
    if (B == 0) return CONE_OK;
    const int S = Lv + Lt;
    const int nkp = (S + (Lv & 1) + 15) & ~15, nt = (S + (Lv & 1) + 127) / 128;  // sized for the layer-0 form (gap row)
    CONE_REQUIRE(enc_attn_tc_supported(Lv, Lt, 256, 8), "enc_attn_tc: unsupported window %d + %d", Lv, Lt);
    CONE_REQUIRE((ldo % 8) == 0, "enc_attn_tc: output rows must be 16-byte aligned");
    const bool indirect = token_qkv != nullptr;
    CONE_REQUIRE(!indirect || (vid_base && txt_base && n_tok > 0), "enc_attn_tc: incomplete row tables");
    CUtensorMap mQ, mQT, mV, mVT;
    // dense: one box of S rows; indirect: Lv frame rows + Lt token rows
    CONE_TRY(make_map_box(&mQ, qkv, rows, 768, 768, indirect ? Lv : S, 64));
    CONE_TRY(make_map_box(&mV, qkv, rows, 768, 768, indirect ? Lv : S, 32));
    mQT = mQ;
    mVT = mV;
    if (indirect) {
        CONE_TRY(make_map_box(&mQT, token_qkv, n_tok, 768, 768, Lt, 64));
        CONE_TRY(make_map_box(&mVT, token_qkv, n_tok, 768, 768, Lt, 32));
    }
    AtParams P{};
    P.out = static_cast<__half*>(o);
    P.ldo = ldo;
    P.vlen = vlen; P.tlen = tlen; P.vid_base = vid_base; P.txt_base = txt_base;
    P.pos = static_cast<const __half*>(posqk16);
    P.B = B; P.Lv = Lv; P.Lt = Lt; P.table_lv = table_lv; P.nkp = nkp; P.ntile = nt; P.indirect = indirect ? 1 : 0;
    P.tpad = indirect ? (Lv & 1) : 0;
    const AtPlan pl = at_plan(nkp, nt);
    static int smem_set = 0;
    if (smem_set < pl.total) {
        CONE_CUDA(cudaFuncSetAttribute(enc_attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, pl.total));
        smem_set = pl.total;
    }
    const unsigned grid = (unsigned)(B < num_sms ? B : num_sms);
    ProfScope ps(s, P_ENC_ATTN, 4.0 * (double)B * 8 * S * S * AT_HD, 8.0 * (double)B * S * 8 * AT_HD);
    enc_attn_tc_kernel<<<grid, AT_THREADS, pl.total, s>>>(mQ, mQT, mV, mVT, P);
    CONE_LAUNCH_CHECK("enc_attn_tc");
    return CONE_OK;
}

}  // namespace cone

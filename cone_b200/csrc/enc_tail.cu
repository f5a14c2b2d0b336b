// Fused tail of a Moment-DETR encoder layer on the tensor cores (CONE_PREC_TC):
//
//     x   = LayerNorm1(res + att . Wo^T + bo)                      cone/transformer.py:239-241 (out_proj, norm1)
//     out = LayerNorm2(x + relu(x . W1^T + b1) . W2^T + b2)        cone/transformer.py:242-245 (linear1/2, norm2)
//
// for a [M, 256] stream of window rows, ONE kernel: the LayerNorm1 output and the [M, 1024] FFN hidden never reach HBM
// (the unfused chain wrote and re-read 7.2 KB per row and layer; this kernel moves 1.5-2.5 KB).
//
// sm_100a design.  Persistent, one CTA per SM, warp-specialised like tc_gemm.cu (warp 0 TMA producer, warp 1 MMA issuer,
// warps 2-17 epilogue), 128 rows per CTA and tile.  With CG = 2 two CTAs of a cluster form a tcgen05 CTA pair
// (`cta_group::2`, M = 256): every weight tile is split between the two CTAs' shared memories, so each SM streams HALF
// of the 1.15 MB of weights per 128 rows from L2 — at one CTA per tile the weight stream alone (10 TB/s over 148 SMs)
// would exceed what L2 delivers.
//
// Data flow of one tile (TMEM: Y = columns [0,256), Hreg = columns [256,512)):
//   GEMM0   Hreg  = res_hi . I + res_lo . I + att . Wo^T   (K = 256)     -- the residual rows ride the same TMA ring as
//           the operands and are added by N = 64 MMAs against a 64 x 64 identity tile (exact: products with 1.0, fp32
//           accumulation), so the epilogue never touches global memory for them
//   epi-0   x = LN1(Hreg + bo): fp16 copy -> X tile in shared memory (A operand of GEMM1, UMMA K-major 128B-swizzle
//           layout written by the epilogue threads), fp32 x + b2 -> Y (tcgen05.st): the fp32 residual of the FFN is
//           the INITIAL VALUE of GEMM2's accumulator and never leaves the SM
//   for each chunk c of 128 hidden columns (double buffered in TMEM and shared memory):
//       GEMM1  Hreg[c & 1] = X . W1[c]^T          epi-h  relu(. + b1) -> fp16 -> H[c & 1] in shared memory
//       GEMM2  Y += H[c & 1] . W2[:, c]^T
//   epi-f   out = LN2(Y) -> fp16 hi (+ fp16 lo = out - hi: the residual stream crosses HBM as hi + lo, the operands of
//           the next GEMMs read hi only) staged in the X / H regions -> TMA stores
// The MMA warp issues GEMM0 of the NEXT tile right after the last GEMM2, so it overlaps the final epilogue.
//
// Gather mode (encoder layer 0, EtParams::gather): the residual rows are the window slices of the per-frame / per-token
// projections (cone/ego4d_mad_dataloader.py:144-159, start_end_collate).  Producer warp A derives the source row of every
// tile row from the window descriptors and fetches each [128 x 64] residual panel with 32 TMA gather4 instructions (four
// arbitrary rows each, one-row boxes, 128-byte swizzle): the panel lands exactly as a tiled box would, nothing else
// changes, and no gathered copy of the window inputs is ever written to HBM.
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include "kernels.h"
#include "tc_gemm.h"
#include "tc_ptx.cuh"

namespace cone {

using namespace ptx;

namespace {

constexpr int ET_THREADS = 640;           // producer A + MMA warp A + 16 epilogue warps + MMA warp B + producer B
constexpr int ET_EPI_WARPS = 16;
constexpr int ET_MMA_B_WARP = 18;
constexpr int ET_PROD_B_WARP = 19;
constexpr int ET_D = 256;               // model width (rows of 256 fp16 = 4 k-blocks of 64)
constexpr int SLOT_BYTES = 16384;       // one [128 x 64] fp16 box, 128-byte swizzle
constexpr int NSLOT = 5;
constexpr int NSLOT_A = 3;             // ring slots of MMA warp A's items (GEMM0, GEMM1); the other 2 carry GEMM2's
constexpr int XT_BYTES = 4 * SLOT_BYTES;   // X tile [128 x 256] fp16
constexpr int HB_BYTES = 2 * SLOT_BYTES;   // one hidden chunk [128 x 128] fp16
constexpr int I64_BYTES = 64 * 128;
constexpr int OFF_X = 0;
constexpr int OFF_H = OFF_X + XT_BYTES;
constexpr int OFF_RING = OFF_H + 2 * HB_BYTES;
constexpr int OFF_I64 = OFF_RING + NSLOT * SLOT_BYTES;
constexpr int OFF_BAR = OFF_I64 + I64_BYTES;
constexpr int N_BAR = 2 * NSLOT + 11;
constexpr int OFF_LN = OFF_BAR + 8 * 32;           // room for 32 barriers
constexpr int OFF_SLOT = OFF_LN + ET_EPI_WARPS * 32 * 8;  // LayerNorm partials [16 warps][32 lanes] float2
constexpr int OFF_PAR = OFF_SLOT + 64;             // bo, ln1_g, ln1_b, b2, ln2_g, ln2_b: 6 x 256 floats
constexpr int ET_SMEM = OFF_PAR + 6 * ET_D * 4;    // no alignment slack: the kernel traps if the window is not 1 KB aligned
static_assert(N_BAR <= 32, "barrier area");
static_assert(ET_SMEM <= 232448, "enc_tail shared memory exceeds 227 KB");

struct EtParams {
    const float *bo, *ln1_g, *ln1_b, *b1, *b2, *ln2_g, *ln2_b;
    float eps;
    float* C32;      // optional fp32 copy of the output rows (saliency head)
    int64_t ldc32;
    int64_t M;
    int nchunk;      // ffn / 128
    int has_lo_in, has_lo_out;
    // gather mode (encoder layer 0): the residual rows are the window slices of the per-frame / per-token projections
    // (cone/ego4d_mad_dataloader.py:144-159 + start_end_collate): tile row m = window b = m / S, row r = m % S comes from source
    // row min(vid_base[b] + r, n_vid - 1) for r < Lv, n_vid + txt_base[b] + (r - Lv) otherwise; tmRhi / tmRlo are maps of the
    // source table [n_vid + n_txt, 256] with one-row boxes and the rows arrive by TMA gather4 — no gathered copy in HBM
    int gather, g_S, g_Lv;
    int64_t g_nvid;
    const int64_t* g_vid_base;
    const int64_t* g_txt_base;
};

__device__ __forceinline__ uint32_t et_idesc(int M, int N) {  // D fp32, A / B fp16 K-major
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <int CG>
__global__ void __launch_bounds__(ET_THREADS, 1)
enc_tail_kernel(const __grid_constant__ CUtensorMap tmAtt, const __grid_constant__ CUtensorMap tmRhi,
                const __grid_constant__ CUtensorMap tmRlo, const __grid_constant__ CUtensorMap tmWo,
                const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
                const __grid_constant__ CUtensorMap tmOhi, const __grid_constant__ CUtensorMap tmOlo, EtParams P) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw;
    if ((smem_u32(smem_raw) & 1023u) != 0) __trap();  // the 128-byte-swizzle atoms need 1 KB alignment
    uint8_t* sX = smem + OFF_X;
    uint8_t* sH = smem + OFF_H;
    uint8_t* sRing = smem + OFF_RING;
    uint8_t* sI = smem + OFF_I64;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* empty = full + NSLOT;
    uint64_t* g0full = empty + NSLOT;
    uint64_t* xready = g0full + 1;
    uint64_t* yfull = xready + 1;
    uint64_t* hfull = yfull + 1;    // [2]
    uint64_t* hready = hfull + 2;   // [2] fp16 chunk in shared memory
    uint64_t* htfree = hready + 2;  // [2] chunk accumulator read out of TMEM
    uint64_t* hsfree = htfree + 2;  // [2] GEMM2 has finished reading the shared-memory chunk buffer
    float2* ln_part = reinterpret_cast<float2*>(smem + OFF_LN);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_SLOT);
    float* sPar = reinterpret_cast<float*>(smem + OFF_PAR);  // epilogue vectors (broadcast LDS instead of global loads)
    const float *s_bo = sPar, *s_g1 = sPar + ET_D, *s_b1 = sPar + 2 * ET_D, *s_b2 = sPar + 3 * ET_D, *s_g2 = sPar + 4 * ET_D,
                *s_bb2 = sPar + 5 * ET_D;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
    const bool leader = rank == 0;
    const int64_t n_super = (P.M + 128 * CG - 1) / (128 * CG);
    const int64_t st_begin = blockIdx.x / CG, st_step = gridDim.x / CG;
    const int nchunk = P.nchunk;
    const int nres = P.has_lo_in ? 8 : 4;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmAtt)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmW1)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmW2)) : "memory");
        for (int i = 0; i < NSLOT; ++i) {
            mbar_init(&full[i], CG);  // one arrive.expect_tx per CTA of the pair
            mbar_init(&empty[i], 1);  // one tcgen05.commit
        }
        mbar_init(g0full, 1);
        mbar_init(yfull, 1);
        mbar_init(xready, ET_EPI_WARPS * CG);  // lane 0 of every epilogue warp of the pair
        for (int i = 0; i < 2; ++i) {
            mbar_init(&hfull[i], 1);
            mbar_init(&hready[i], (ET_EPI_WARPS / 2) * CG);  // the 8 warps of the group that owns this chunk buffer
            mbar_init(&htfree[i], (ET_EPI_WARPS / 2) * CG);
            mbar_init(&hsfree[i], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // the whole TMEM: Y (256 columns) + two hidden-chunk accumulators / the GEMM0 accumulator (256)
        if (CG == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        }
    }
    for (int i = threadIdx.x; i < ET_D; i += ET_THREADS) {
        sPar[i] = P.bo[i];
        sPar[ET_D + i] = P.ln1_g[i];
        sPar[2 * ET_D + i] = P.ln1_b[i];
        sPar[3 * ET_D + i] = P.b2[i];
        sPar[4 * ET_D + i] = P.ln2_g[i];
        sPar[5 * ET_D + i] = P.ln2_b[i];
    }
    // identity tile of the residual MMAs: this CTA's 64 / CG rows of I64 (row n holds a single 1.0 at k = n)
    for (int i = threadIdx.x; i < I64_BYTES / 16; i += ET_THREADS) reinterpret_cast<uint4*>(sI)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    if (threadIdx.x < 64 / CG) {
        const int nl = threadIdx.x, n = (64 / CG) * (int)rank + nl;
        *reinterpret_cast<__half*>(sI + sw128(nl, n >> 3) + (n & 7) * 2) = __float2half_rn(1.0f);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();  // barrier inits and TMEM allocation of BOTH CTAs before any cross-CTA traffic
    tc_fence_after();
    // All 512 columns are allocated by the only CTA on this SM: the allocation starts at lane 0 / column 0.  Using the
    // literal keeps every TMEM address a compile-time constant (uniform registers in the MMA warp).
    if (*tmem_slot != 0u) __trap();
    constexpr uint32_t tmem_base = 0u;
    constexpr uint32_t tmemY = tmem_base, tmemH = tmem_base + 256;

    if (warp == 0 && P.gather) {  // ---------------------------------------- producer A with gathered residual rows (whole warp)
        int slot = 0;
        uint32_t ph = 0;
        auto advance = [&]() {
            if (++slot == NSLOT_A) {
                slot = 0;
                ph ^= 1;
            }
        };
        auto emit = [&](const CUtensorMap* map, int c0, int c1) {  // one [128 x 64] box, issued by lane 0
            mbar_wait(&empty[slot], ph ^ 1);
            if (lane == 0) {
                if (CG == 1) {
                    mbar_expect_tx(&full[slot], SLOT_BYTES);
                    tma_load_2d(sRing + slot * SLOT_BYTES, map, &full[slot], c0, c1);
                } else {
                    mbar_expect_tx_leader(&full[slot], SLOT_BYTES);
                    tma_load_2d_pair(sRing + slot * SLOT_BYTES, map, &full[slot], c0, c1);
                }
            }
            __syncwarp();
            advance();
        };
        auto emit_rows = [&](const CUtensorMap* map, int kb, const int* idx) {  // lane l: tile rows 4 l .. 4 l + 3 of the box
            mbar_wait(&empty[slot], ph ^ 1);
            if (lane == 0) {
                if (CG == 1) mbar_expect_tx(&full[slot], SLOT_BYTES);
                else mbar_expect_tx_leader(&full[slot], SLOT_BYTES);
            }
            __syncwarp();
            tma_gather4<CG>(sRing + slot * SLOT_BYTES + lane * 512, map, &full[slot], kb * 64, idx[0], idx[1], idx[2], idx[3]);
            __syncwarp();
            advance();
        };
        auto source_rows = [&](int m0, int* idx) {  // lane l: source rows of tile rows 4 l .. 4 l + 3
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int64_t m = (int64_t)m0 + 4 * lane + j;
                if (m >= P.M) m = P.M - 1;  // rows past the end are clipped by the output stores
                const int64_t b = m / P.g_S;
                const int r = (int)(m - b * P.g_S);
                int64_t src;
                if (r < P.g_Lv) {
                    src = P.g_vid_base[b] + r;
                    src = src < P.g_nvid ? src : P.g_nvid - 1;
                } else {
                    src = P.g_nvid + P.g_txt_base[b] + (r - P.g_Lv);
                }
                idx[j] = (int)src;
            }
        };
        auto emit_gemm0 = [&](int m0, const int* idx) {
            // the residual panels go FIRST: they are the slow loads (128 row segments each), and the first ring slots of a
            // tile are filled while the previous tile's last GEMM1 chunks are still being multiplied
            for (int kb = 0; kb < 4; ++kb) emit_rows(&tmRhi, kb, idx);
            if (P.has_lo_in)
                for (int kb = 0; kb < 4; ++kb) emit_rows(&tmRlo, kb, idx);
            for (int kb = 0; kb < 4; ++kb) {
                emit(&tmAtt, kb * 64, m0);
                if (CG == 1) {
                    emit(&tmWo, kb * 64, 0);
                    emit(&tmWo, kb * 64, 128);
                } else {
                    emit(&tmWo, kb * 64, 128 * (int)rank);
                }
            }
        };
        auto emit_g1 = [&](int c) {
            if (CG == 1) {
                for (int kb = 0; kb < 4; ++kb) emit(&tmW1, kb * 64, c * 128);
            } else {
                for (int kp = 0; kp < 2; ++kp) {
                    mbar_wait(&empty[slot], ph ^ 1);
                    if (lane == 0) {
                        mbar_expect_tx_leader(&full[slot], SLOT_BYTES);
                        tma_load_2d_pair(sRing + slot * SLOT_BYTES, &tmW1, &full[slot], (2 * kp) * 64, c * 128 + 64 * (int)rank);
                        tma_load_2d_pair(sRing + slot * SLOT_BYTES + SLOT_BYTES / 2, &tmW1, &full[slot], (2 * kp + 1) * 64,
                                         c * 128 + 64 * (int)rank);
                    }
                    __syncwarp();
                    advance();
                }
            }
        };
        int idx[4];
        if (st_begin < n_super) {
            const int m_first = (int)((st_begin * CG + rank) * 128);
            source_rows(m_first, idx);
            emit_gemm0(m_first, idx);
        }
        for (int64_t st = st_begin; st < n_super; st += st_step) {
            const bool has_next = st + st_step < n_super;
            const int m_next = (int)(((st + st_step) * CG + rank) * 128);
            if (has_next) {
                // (an L2 prefetch of the next tile's gathered rows was measured: +0.7 ms per launch — the TMA unit's rate of
                // ~10 cycles per 128-byte row segment is what the gather costs, not L2 misses)
                source_rows(m_next, idx);
                if (lane == 0)
                    for (int kb = 0; kb < 4; ++kb) tma_prefetch_l2_2d(&tmAtt, kb * 64, m_next);
            }
            for (int c = 0; c < nchunk; ++c) emit_g1(c);
            if (has_next) emit_gemm0(m_next, idx);
        }
    } else if (warp == 0 || warp == ET_PROD_B_WARP) {
        if (lane == 0) {  // ------------------------------------------------------------------------------ TMA producers
            // Two sub-rings with their own cursors: slots [0, NSLOT_A) carry the items of MMA warp A (GEMM0, GEMM1), slots
            // [NSLOT_A, NSLOT) those of warp B (GEMM2).  (With ONE ring shared by two consumers a consumer that skips the
            // other's items can get a whole lap ahead of an item that has not landed yet, and a parity wait on that slot
            // then passes at once: the first two-issuer build failed exactly so.)
            int slot[2] = {0, NSLOT_A};
            uint32_t ph[2] = {0, 0};
            auto advance = [&](int r) {
                if (++slot[r] == (r == 0 ? NSLOT_A : NSLOT)) {
                    slot[r] = (r == 0 ? 0 : NSLOT_A);
                    ph[r] ^= 1;
                }
            };
            auto emit = [&](int r, const CUtensorMap* map, int c0, int c1, uint32_t bytes) {
                const int sl = slot[r];
                mbar_wait(&empty[sl], ph[r] ^ 1);
                if (CG == 1) {
                    mbar_expect_tx(&full[sl], bytes);
                    tma_load_2d(sRing + sl * SLOT_BYTES, map, &full[sl], c0, c1);
                } else {
                    mbar_expect_tx_leader(&full[sl], bytes);
                    tma_load_2d_pair(sRing + sl * SLOT_BYTES, map, &full[sl], c0, c1);
                }
                advance(r);
            };
            auto emit_gemm0 = [&](int m0) {  // (same item order as the gather-mode producer: residual panels first)
                for (int kb = 0; kb < 4; ++kb) emit(0, &tmRhi, kb * 64, m0, SLOT_BYTES);
                if (P.has_lo_in)
                    for (int kb = 0; kb < 4; ++kb) emit(0, &tmRlo, kb * 64, m0, SLOT_BYTES);
                for (int kb = 0; kb < 4; ++kb) {
                    emit(0, &tmAtt, kb * 64, m0, SLOT_BYTES);
                    if (CG == 1) {
                        emit(0, &tmWo, kb * 64, 0, SLOT_BYTES);
                        emit(0, &tmWo, kb * 64, 128, SLOT_BYTES);
                    } else {
                        emit(0, &tmWo, kb * 64, 128 * (int)rank, SLOT_BYTES);
                    }
                }
            };
            auto emit_g1 = [&](int c) {  // CG = 2: this CTA's 64 rows of the chunk, TWO k-blocks per 16 KB slot
                if (CG == 1) {
                    for (int kb = 0; kb < 4; ++kb) emit(0, &tmW1, kb * 64, c * 128, SLOT_BYTES);
                } else {
                    for (int kp = 0; kp < 2; ++kp) {
                        const int sl = slot[0];
                        mbar_wait(&empty[sl], ph[0] ^ 1);
                        mbar_expect_tx_leader(&full[sl], SLOT_BYTES);
                        tma_load_2d_pair(sRing + sl * SLOT_BYTES, &tmW1, &full[sl], (2 * kp) * 64, c * 128 + 64 * (int)rank);
                        tma_load_2d_pair(sRing + sl * SLOT_BYTES + SLOT_BYTES / 2, &tmW1, &full[sl], (2 * kp + 1) * 64,
                                         c * 128 + 64 * (int)rank);
                        advance(0);
                    }
                }
            };
            auto emit_g2 = [&](int c) {
                for (int kb = 0; kb < 2; ++kb) {
                    if (CG == 1) {
                        emit(1, &tmW2, c * 128 + kb * 64, 0, SLOT_BYTES);
                        emit(1, &tmW2, c * 128 + kb * 64, 128, SLOT_BYTES);
                    } else {
                        emit(1, &tmW2, c * 128 + kb * 64, 128 * (int)rank, SLOT_BYTES);
                    }
                }
            };
            auto prefetch_rows = [&](int m0) {  // next tile's activation rows -> L2, a whole tile ahead of their TMA loads
                for (int kb = 0; kb < 4; ++kb) {
                    tma_prefetch_l2_2d(&tmAtt, kb * 64, m0);
                    tma_prefetch_l2_2d(&tmRhi, kb * 64, m0);
                    if (P.has_lo_in) tma_prefetch_l2_2d(&tmRlo, kb * 64, m0);
                }
            };
            if (warp == 0) {  // producer of sub-ring A: GEMM0 and GEMM1 items, in MMA warp A's order
                if (st_begin < n_super) emit_gemm0((int)((st_begin * CG + rank) * 128));
                for (int64_t st = st_begin; st < n_super; st += st_step) {
                    const bool has_next = st + st_step < n_super;
                    const int m_next = (int)(((st + st_step) * CG + rank) * 128);
                    if (has_next) prefetch_rows(m_next);
                    for (int c = 0; c < nchunk; ++c) emit_g1(c);
                    if (has_next) emit_gemm0(m_next);
                }
            } else {  // producer of sub-ring B: GEMM2 items
                for (int64_t st = st_begin; st < n_super; st += st_step)
                    for (int c = 0; c < nchunk; ++c) emit_g2(c);
            }
        }
    } else if (warp == 1 || warp == ET_MMA_B_WARP) {
        if (leader) {  // ------------------------------------------------------------------------------------ MMA issuers
            // TWO issuing warps.  ncu on the single-issuer version: operands always landed (0 retries on the ring's `full`
            // barriers), epilogue hand-offs always ready (0 retries on htfree / hready), tensor pipe 51 % busy — the issuing
            // warp itself (barrier try_wait ~90 cycles even when complete, commit, descriptor set-up: ~190 warp-operations
            // per tile) was the critical path.  Warp A issues GEMM0 and the GEMM1 chunks, warp B the GEMM2 chunks; they
            // never touch the same accumulator and every dependency between them goes through the epilogue's barriers.
            // Each consumes its own sub-ring of TMA slots (see the producer).
            // The WHOLE warp runs the role (uniform control flow, waits included); one elected lane issues the tcgen05
            // instructions, so that descriptors stay in uniform registers.
            const bool is_a = (warp == 1);
            const uint32_t id256 = et_idesc(128 * CG, 256), id128 = et_idesc(128 * CG, 128), id64 = et_idesc(128 * CG, 64);
            const int slot_lo = is_a ? 0 : NSLOT_A, slot_hi = is_a ? NSLOT_A : NSLOT;  // this warp's sub-ring
            int slot = slot_lo;
            uint32_t ph = 0;
            // descriptor `lo` words (start address >> 4): byte offsets below are added as (bytes >> 4)
            const uint32_t ring_addr = desc_lo_sw128(smem_u32(sRing)), x_addr = desc_lo_sw128(smem_u32(sX)),
                           h_addr = desc_lo_sw128(smem_u32(sH)), i_addr = desc_lo_sw128(smem_u32(sI));
            constexpr uint32_t SLOT16 = SLOT_BYTES >> 4, HB16 = HB_BYTES >> 4;
            auto take = [&]() -> int {  // next item of this warp's sub-ring has landed (in both CTAs of the pair)
                mbar_wait(&full[slot], ph);
                tc_fence_after();
                const int s = slot;
                if (++slot == slot_hi) {
                    slot = slot_lo;
                    ph ^= 1;
                }
                return s;
            };
            auto mma4 = [&](uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, bool acc_first) {
                // the four K = 16 steps of one 64-column k-block: +32 bytes = +2 in the start-address field
                if (elect_one_sync()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t acc = (acc_first || k > 0) ? 1u : 0u;
                        if (CG == 1) umma_f16_lo(d, a_lo + 2 * k, b_lo + 2 * k, idesc, acc);
                        else umma_f16_pair_lo(d, a_lo + 2 * k, b_lo + 2 * k, idesc, acc);
                    }
                }
                __syncwarp();
            };
            auto commit = [&](uint64_t* bar) {
                if (elect_one_sync()) {
                    if (CG == 1) umma_commit(bar);
                    else umma_commit_pair(bar);
                }
                __syncwarp();
            };
            auto gemm0 = [&]() {
                for (int j = 0; j < nres; ++j) {  // res_hi (+ res_lo): 64 columns at a time against the identity tile; they
                    const int ia = take();        // open the accumulation (the first product of each block overwrites)
                    mma4(tmemH + 64 * (j & 3), ring_addr + ia * SLOT16, i_addr, id64, j >= 4);
                    commit(&empty[ia]);
                }
                for (int kb = 0; kb < 4; ++kb) {  // + att . Wo^T
                    const int ia = take(), ib0 = take(), ib1 = (CG == 1) ? take() : 0;
                    const uint32_t a = ring_addr + ia * SLOT16, b0 = ring_addr + ib0 * SLOT16, b1 = ring_addr + ib1 * SLOT16;
                    if (CG == 1) {
                        mma4(tmemH, a, b0, id128, true);
                        mma4(tmemH + 128, a, b1, id128, true);
                    } else {
                        mma4(tmemH, a, b0, id256, true);
                    }
                    commit(&empty[ia]);
                    commit(&empty[ib0]);
                    if (CG == 1) commit(&empty[ib1]);
                }
                commit(g0full);
            };
            auto gemm1 = [&](int c) {  // hidden chunk c: Hreg[c & 1] = X . W1[c]^T
                const uint32_t d = tmemH + 128 * (c & 1);
                if (CG == 1) {
                    for (int kb = 0; kb < 4; ++kb) {
                        const int ib = take();
                        mma4(d, x_addr + kb * SLOT16, ring_addr + ib * SLOT16, id128, kb > 0);
                        commit(&empty[ib]);
                    }
                } else {
                    for (int kp = 0; kp < 2; ++kp) {  // one slot = k-blocks 2 kp, 2 kp + 1 of this CTA's 64 rows
                        const int ib = take();
                        mma4(d, x_addr + (2 * kp) * SLOT16, ring_addr + ib * SLOT16, id128, kp > 0);
                        mma4(d, x_addr + (2 * kp + 1) * SLOT16, ring_addr + ib * SLOT16 + SLOT16 / 2, id128, true);
                        commit(&empty[ib]);
                    }
                }
                commit(&hfull[c & 1]);
            };
            auto gemm2 = [&](int c) {  // Y += H[c & 1] . W2[:, c]^T
                for (int kb = 0; kb < 2; ++kb) {
                    const int ib0 = take(), ib1 = (CG == 1) ? take() : 0;
                    const uint32_t a = h_addr + (c & 1) * HB16 + kb * SLOT16, b0 = ring_addr + ib0 * SLOT16, b1 = ring_addr + ib1 * SLOT16;
                    if (CG == 1) {
                        mma4(tmemY, a, b0, id128, true);
                        mma4(tmemY + 128, a, b1, id128, true);
                    } else {
                        mma4(tmemY, a, b0, id256, true);
                    }
                    commit(&empty[ib0]);
                    if (CG == 1) commit(&empty[ib1]);
                }
            };
            uint32_t p = 0, bph[2] = {0, 0};  // warp A: htfree phases, warp B: hready phases
            if (st_begin < n_super && is_a) gemm0();
            for (int64_t st = st_begin; st < n_super; st += st_step) {
                const bool has_next = st + st_step < n_super;
                if (is_a) {
                    mbar_wait(xready, p);
                    tc_fence_after();
                    gemm1(0);
                    if (nchunk > 1) gemm1(1);
                    for (int c = 0; c < nchunk; ++c) {
                        const int b = c & 1;
                        // chunk c's accumulator is in the epilogue's registers: GEMM1 of chunk c + 2 may overwrite it
                        mbar_wait(&htfree[b], bph[b]);
                        bph[b] ^= 1;
                        tc_fence_after();
                        if (c + 2 < nchunk) gemm1(c + 2);
                    }
                    if (has_next) gemm0();  // overlaps the final epilogue of this tile
                } else {
                    for (int c = 0; c < nchunk; ++c) {
                        const int b = c & 1;
                        mbar_wait(&hready[b], bph[b]);  // the fp16 chunk is in shared memory (and Y holds this tile's residual)
                        bph[b] ^= 1;
                        tc_fence_after();
                        gemm2(c);
                        if (c + 2 < nchunk) commit(&hsfree[b]);  // the epilogue writes chunk c + 2 into the same buffer
                    }
                    commit(yfull);
                }
                p ^= 1;
            }
        }
    } else if (warp >= 2 && warp < 2 + ET_EPI_WARPS) {  // ------------------------------------------------ epilogue warps
        // 16 warps: four per TMEM lane quarter, each owning a quarter of the columns (64 of the 256-wide row in the
        // LayerNorm epilogues, 32 of a 128-wide hidden chunk).  These phases are dependent chains (TMEM load -> math ->
        // convert -> store) that two warps per scheduler cannot keep busy: with 8 warps ncu showed 15 500 cycles per tile
        // in the two LayerNorm epilogues alone, during which the tensor pipe only has GEMM0 to do.
        const int ew = warp - 2;       // 0..15
        const int quarter = warp & 3;  // TMEM lane quarter this warp may access
        const int part = ew >> 2;      // which quarter of the columns
        const int row = quarter * 32 + lane;
        const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
        uint32_t p = 0, hfph[2] = {0, 0}, hsph[2] = {0, 0};
        auto arrive = [&](uint64_t* bar) {
            if (CG == 1) mbar_arrive(bar);
            else mbar_arrive_leader(bar);
        };
        // LayerNorm statistics of this lane's row: 64 columns here, 64 in each of the three partner warps
        auto row_stats = [&](float s1, float s2, float& mean, float& rstd) {
            ln_part[ew * 32 + lane] = make_float2(s1, s2);
            asm volatile("bar.sync %0, 128;" ::"r"(1 + quarter) : "memory");
            s1 = 0.f;
            s2 = 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float2 o = ln_part[((ew & 3) + 4 * q) * 32 + lane];
                s1 += o.x;
                s2 += o.y;
            }
            asm volatile("bar.sync %0, 128;" ::"r"(1 + quarter) : "memory");
            mean = s1 * (1.f / ET_D);
            rstd = rsqrtf(fmaxf(s2 * (1.f / ET_D) - mean * mean, 0.f) + P.eps);
        };
        // sum and sum of squares of 64 values, four independent chains
        auto stats64 = [&](const float* x, float& s1, float& s2) {
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll
            for (int j = 0; j < 64; j += 4) {
                a0 += x[j]; a1 += x[j + 1]; a2 += x[j + 2]; a3 += x[j + 3];
                q0 = fmaf(x[j], x[j], q0); q1 = fmaf(x[j + 1], x[j + 1], q1);
                q2 = fmaf(x[j + 2], x[j + 2], q2); q3 = fmaf(x[j + 3], x[j + 3], q3);
            }
            s1 = (a0 + a1) + (a2 + a3);
            s2 = (q0 + q1) + (q2 + q3);
        };
        for (int64_t st = st_begin; st < n_super; st += st_step) {
            const int m0 = (int)((st * CG + rank) * 128);
            const int col = 64 * part;  // first of this warp's 64 columns in the LayerNorm epilogues
            // ---- epi-0: x = LN1(att.Wo^T + res + bo) -> X tile (fp16), Y (fp32, + b2); one TMEM read, values in registers
            mbar_wait(g0full, p);
            tc_fence_after();
            if (lane == 0) tma_store_wait_read<0>();  // the staged output of the previous tile has left X / H
            __syncwarp();
            {
                float x[64];
                tmem_ld_32x64(tmemH + lane_base + col, x);
                tmem_ld_wait();
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    const float4 b = reinterpret_cast<const float4*>(s_bo + col)[q];
                    x[4 * q] += b.x; x[4 * q + 1] += b.y; x[4 * q + 2] += b.z; x[4 * q + 3] += b.w;
                }
                float s1, s2, mean, rstd;
                stats64(x, s1, s2);
                row_stats(s1, s2, mean, rstd);
                const float nm = -mean * rstd;
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    const float4 g = reinterpret_cast<const float4*>(s_g1 + col)[q];
                    const float4 b = reinterpret_cast<const float4*>(s_b1 + col)[q];
                    x[4 * q] = fmaf(fmaf(x[4 * q], rstd, nm), g.x, b.x);
                    x[4 * q + 1] = fmaf(fmaf(x[4 * q + 1], rstd, nm), g.y, b.y);
                    x[4 * q + 2] = fmaf(fmaf(x[4 * q + 2], rstd, nm), g.z, b.z);
                    x[4 * q + 3] = fmaf(fmaf(x[4 * q + 3], rstd, nm), g.w, b.w);
                }
                uint8_t* xk = sX + part * SLOT_BYTES;  // k-block `part` of the X tile
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    uint4 v;
                    v.x = pack_h2(x[8 * u], x[8 * u + 1]);
                    v.y = pack_h2(x[8 * u + 2], x[8 * u + 3]);
                    v.z = pack_h2(x[8 * u + 4], x[8 * u + 5]);
                    v.w = pack_h2(x[8 * u + 6], x[8 * u + 7]);
                    *reinterpret_cast<uint4*>(xk + sw128(row, u)) = v;
                }
#pragma unroll
                for (int q = 0; q < 16; ++q) {  // Y starts as the fp32 residual of the FFN + linear2's bias
                    const float4 b = reinterpret_cast<const float4*>(s_b2 + col)[q];
                    x[4 * q] += b.x; x[4 * q + 1] += b.y; x[4 * q + 2] += b.z; x[4 * q + 3] += b.w;
                }
                tmem_st_32x32(tmemY + lane_base + col, x);
                tmem_st_32x32(tmemY + lane_base + col + 32, x + 32);
                tmem_st_wait();
                fence_async_smem();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) arrive(xready);
            }
            // ---- epi-h: hidden chunks.  The 16 warps form two groups of 8 (two per TMEM lane quarter, 64 columns each):
            // group g owns chunk buffer g and converts the chunks c = g, g + 2, ...; the two groups work on consecutive
            // chunks CONCURRENTLY, so the hand-off chain of one chunk (commit -> wake-up -> TMEM load -> convert -> store ->
            // proxy fence -> arrive -> MMA issue) may take two chunk times instead of one before the tensor pipe waits.
            // Two hand-offs per chunk: `htfree` as soon as the accumulator is in registers (GEMM1 of chunk c + 2 may overwrite
            // it), `hready` once the fp16 chunk is in shared memory (GEMM2 of this chunk may start).
            {
                const int grp = ew >> 3;        // chunk buffer / chunk parity of this warp
                const int hp = (ew >> 2) & 1;   // which 64 of the chunk's 128 columns
                for (int c = grp; c < nchunk; c += 2) {
                    const int b = grp;
                    mbar_wait(&hfull[b], hfph[b]);
                    hfph[b] ^= 1;
                    tc_fence_after();
                    float x[64];
                    tmem_ld_32x64(tmemH + lane_base + 128 * b + 64 * hp, x);
                    const float4* b4 = reinterpret_cast<const float4*>(P.b1 + c * 128 + 64 * hp);
                    float4 bias[8];  // first half of the bias, issued while the TMEM load is in flight
#pragma unroll
                    for (int q = 0; q < 8; ++q) bias[q] = __ldg(b4 + q);
                    tmem_ld_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) arrive(&htfree[b]);
                    uint32_t hv[32];  // relu(x + b1) as packed fp16 pairs
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 bb = bias[q];
                        hv[2 * q] = pack_h2(fmaxf(x[4 * q] + bb.x, 0.f), fmaxf(x[4 * q + 1] + bb.y, 0.f));
                        hv[2 * q + 1] = pack_h2(fmaxf(x[4 * q + 2] + bb.z, 0.f), fmaxf(x[4 * q + 3] + bb.w, 0.f));
                    }
#pragma unroll
                    for (int q = 0; q < 8; ++q) bias[q] = __ldg(b4 + 8 + q);
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 bb = bias[q];
                        hv[16 + 2 * q] = pack_h2(fmaxf(x[32 + 4 * q] + bb.x, 0.f), fmaxf(x[32 + 4 * q + 1] + bb.y, 0.f));
                        hv[16 + 2 * q + 1] = pack_h2(fmaxf(x[32 + 4 * q + 2] + bb.z, 0.f), fmaxf(x[32 + 4 * q + 3] + bb.w, 0.f));
                    }
                    if (c >= 2) {  // GEMM2 of chunk c - 2 has finished reading this shared-memory buffer
                        mbar_wait(&hsfree[b], hsph[b]);
                        hsph[b] ^= 1;
                    }
                    uint8_t* hk = sH + b * HB_BYTES + hp * SLOT_BYTES;  // k-block hp of chunk buffer b
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        *reinterpret_cast<uint4*>(hk + sw128(row, u)) = make_uint4(hv[4 * u], hv[4 * u + 1], hv[4 * u + 2], hv[4 * u + 3]);
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) arrive(&hready[b]);
                }
            }
            // ---- epi-f: out = LN2(Y) -> hi (+ lo) staged in X / H -> TMA stores
            mbar_wait(yfull, p);
            tc_fence_after();
            {
                float x[64];
                tmem_ld_32x64(tmemY + lane_base + col, x);
                tmem_ld_wait();
                tc_fence_before();  // Y is in registers: this warp's next writes to it are epi-0 of the next tile
                float s1, s2, mean, rstd;
                stats64(x, s1, s2);
                row_stats(s1, s2, mean, rstd);
                const float nm = -mean * rstd;
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    const float4 g = reinterpret_cast<const float4*>(s_g2 + col)[q];
                    const float4 b = reinterpret_cast<const float4*>(s_bb2 + col)[q];
                    x[4 * q] = fmaf(fmaf(x[4 * q], rstd, nm), g.x, b.x);
                    x[4 * q + 1] = fmaf(fmaf(x[4 * q + 1], rstd, nm), g.y, b.y);
                    x[4 * q + 2] = fmaf(fmaf(x[4 * q + 2], rstd, nm), g.z, b.z);
                    x[4 * q + 3] = fmaf(fmaf(x[4 * q + 3], rstd, nm), g.w, b.w);
                }
                const int64_t grow = (int64_t)m0 + row;
                if (P.C32 != nullptr && grow < P.M) {
                    float4* o4 = reinterpret_cast<float4*>(P.C32 + grow * P.ldc32 + col);
#pragma unroll
                    for (int q = 0; q < 16; ++q) o4[q] = make_float4(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
                }
                const uint32_t off = (uint32_t)part * SLOT_BYTES;  // k-block `part` of the staged [128 x 256] tile
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    float lo[8];
                    uint4 v;
                    uint32_t* vw = &v.x;
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const __half2 h = __floats2half2_rn(x[8 * u + 2 * e], x[8 * u + 2 * e + 1]);
                        const float2 f = __half22float2(h);
                        lo[2 * e] = x[8 * u + 2 * e] - f.x;
                        lo[2 * e + 1] = x[8 * u + 2 * e + 1] - f.y;
                        vw[e] = *reinterpret_cast<const uint32_t*>(&h);
                    }
                    *reinterpret_cast<uint4*>(sX + off + sw128(row, u)) = v;
                    if (P.has_lo_out) {
                        uint4 w;
                        w.x = pack_h2(lo[0], lo[1]);
                        w.y = pack_h2(lo[2], lo[3]);
                        w.z = pack_h2(lo[4], lo[5]);
                        w.w = pack_h2(lo[6], lo[7]);
                        *reinterpret_cast<uint4*>(sH + off + sw128(row, u)) = w;
                    }
                }
                fence_async_smem();
                __syncwarp();
                if (lane == 0) {
                    const uint32_t boff = off + (uint32_t)quarter * 4096u;
                    tma_store_2d(&tmOhi, sX + boff, part * 64, m0 + quarter * 32);
                    if (P.has_lo_out) tma_store_2d(&tmOlo, sH + boff, part * 64, m0 + quarter * 32);
                    tma_store_commit();
                }
            }
            p ^= 1;
        }
        if (lane == 0) tma_store_wait_read<0>();  // shared memory must outlive the last bulk stores
    }
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();  // the peer's MMAs read this CTA's shared memory, its commits arrive here
    if (warp == 1) {
        tc_fence_after();
        if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

}  // namespace

int enc_tail_supported(int d, int ffn) { return d == ET_D && ffn >= 128 && (ffn % 128) == 0; }

int enc_tail_run(TcWeights* t, const EncTailArgs& a, cudaStream_t s) {
    CONE_REQUIRE(t != nullptr, "enc_tail: tensor-core weights not initialised");
    CONE_REQUIRE(enc_tail_supported(a.d, a.ffn), "enc_tail: unsupported width d=%d ffn=%d", a.d, a.ffn);
    CONE_REQUIRE(a.M >= 1 && a.M < ((int64_t)1 << 31) - 512, "enc_tail: bad row count %lld", (long long)a.M);
    CONE_REQUIRE(a.att16 && a.res_hi && a.out_hi, "enc_tail: null argument");
    const bool gather = a.g_vid_base != nullptr;
    CONE_REQUIRE(!gather || (a.g_txt_base && a.g_S > 0 && a.g_Lv > 0 && a.g_Lv <= a.g_S && a.g_nvid > 0 && a.g_nsrc > a.g_nvid &&
                             a.g_nsrc < ((int64_t)1 << 31) && a.ldr == a.d),
                 "enc_tail: incomplete gather description");
    CONE_REQUIRE((a.lda % 8) == 0 && (a.ldr % 8) == 0 && (a.ldo % 8) == 0, "enc_tail: row pitches must be multiples of 8");
    int cg = a.cta_group;
    if (cg == 0) {
        static int env_cg = -1;
        if (env_cg < 0) {
            const char* e = getenv("CONE_ENC_TAIL_CG");
            env_cg = (e && e[0] == '1') ? 1 : 2;
        }
        cg = env_cg;
    }
    CONE_REQUIRE(cg == 1 || cg == 2, "enc_tail: cta_group must be 1 or 2");
    const uint16_t *wo = nullptr, *w1 = nullptr, *w2 = nullptr;
    CONE_TRY(tc_weight_f16(t, a.Wo, a.d, a.d, s, &wo));
    CONE_TRY(tc_weight_f16(t, a.W1, a.ffn, a.d, s, &w1));
    CONE_TRY(tc_weight_f16(t, a.W2, a.d, a.ffn, s, &w2));
    CUtensorMap mAtt, mRhi, mRlo, mWo, mW1, mW2, mOhi, mOlo;
    CONE_TRY(tc_make_map(&mAtt, a.att16, false, a.M, a.d, a.lda, 64, 128));
    // residual rows: dense [M, d] tiles, or (gather mode) one-row boxes of the source table [g_nsrc, d] for TMA gather4
    CONE_TRY(tc_make_map(&mRhi, a.res_hi, false, gather ? a.g_nsrc : a.M, a.d, a.ldr, 64, gather ? 1 : 128));
    mRlo = mRhi;
    if (a.res_lo) CONE_TRY(tc_make_map(&mRlo, a.res_lo, false, gather ? a.g_nsrc : a.M, a.d, a.ldr, 64, gather ? 1 : 128));
    CONE_TRY(tc_make_map(&mWo, wo, false, a.d, a.d, a.d, 64, 128));
    CONE_TRY(tc_make_map(&mW1, w1, false, a.ffn, a.d, a.d, 64, 128 / cg));
    CONE_TRY(tc_make_map(&mW2, w2, false, a.d, a.ffn, a.ffn, 64, 128));
    CONE_TRY(tc_make_map(&mOhi, a.out_hi, false, a.M, a.d, a.ldo, 64, 32));
    mOlo = mOhi;
    if (a.out_lo) CONE_TRY(tc_make_map(&mOlo, a.out_lo, false, a.M, a.d, a.ldo, 64, 32));
    EtParams P{};
    P.bo = a.bo; P.ln1_g = a.ln1_g; P.ln1_b = a.ln1_b; P.b1 = a.b1; P.b2 = a.b2; P.ln2_g = a.ln2_g; P.ln2_b = a.ln2_b;
    P.eps = 1e-5f;
    P.C32 = a.C32; P.ldc32 = a.ldc32;
    P.M = a.M;
    P.nchunk = a.ffn / 128;
    P.has_lo_in = a.res_lo != nullptr;
    P.has_lo_out = a.out_lo != nullptr;
    P.gather = gather ? 1 : 0;
    P.g_S = a.g_S; P.g_Lv = a.g_Lv; P.g_nvid = a.g_nvid; P.g_vid_base = a.g_vid_base; P.g_txt_base = a.g_txt_base;
    const int num_sms = tc_num_sms(t);
    const int64_t n_super = cdiv64(a.M, 128 * cg);
    int64_t clusters = num_sms / cg;
    if (clusters > n_super) clusters = n_super;
    const unsigned grid = (unsigned)(clusters * cg);
    const double m = (double)a.M;
    ProfScope ps(s, gather ? P_ENC_TAIL_G : P_ENC_TAIL, 2.0 * m * ((double)a.d * a.d + 2.0 * a.d * a.ffn),
                 2.0 * m * a.d * (2.0 + (a.res_lo ? 1.0 : 0.0) + 1.0 + (a.out_lo ? 1.0 : 0.0)) + (a.C32 ? 4.0 * m * a.d : 0.0));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(ET_THREADS);
    cfg.dynamicSmemBytes = ET_SMEM;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cg;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (cg == 1) {
        static bool set1 = false;
        if (!set1) {
            CONE_CUDA(cudaFuncSetAttribute(enc_tail_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, ET_SMEM));
            set1 = true;
        }
        CONE_CUDA(cudaLaunchKernelEx(&cfg, enc_tail_kernel<1>, mAtt, mRhi, mRlo, mWo, mW1, mW2, mOhi, mOlo, P));
    } else {
        static bool set2 = false;
        if (!set2) {
            CONE_CUDA(cudaFuncSetAttribute(enc_tail_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, ET_SMEM));
            set2 = true;
        }
        CONE_CUDA(cudaLaunchKernelEx(&cfg, enc_tail_kernel<2>, mAtt, mRhi, mRlo, mWo, mW1, mW2, mOhi, mOlo, P));
    }
    CONE_LAUNCH_CHECK("enc_tail");
    return CONE_OK;
}

}  // namespace cone

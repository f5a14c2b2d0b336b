// Tensor-core GEMM for the dense projections (CONE_PREC_TC): C = epi(A * W^T), fp16 operands, fp32 accumulate.
//
// sm_100a design: persistent CTAs (one per SM), warp-specialised:
//   warp 0   TMA producer   cp.async.bulk.tensor 2-D boxes of A [128 x 64] and W [BN x 64] (fp16, 128-byte swizzle)
//                           into a 4-stage shared-memory ring, completion on mbarriers
//   warp 1   MMA issuer     one thread issues tcgen05.mma (cta_group::1, kind::f16, M=128, N=BN, K=16) from
//                           shared-memory descriptors; accumulators live in TMEM (2 x BN fp32 columns, double
//                           buffered so the epilogue of tile i overlaps the MMAs of tile i+1)
//   warps 2-5 epilogue      tcgen05.ld TMEM -> registers (one accumulator row per thread), then fused
//                           bias / residual / ReLU / LayerNorm(N = 256) and fp32 and/or fp16 stores
// fp16 (11 significant bits) rather than bf16: the reference comparison needs 1e-3 on spans and scores, which
// bf16 operands miss by 3-5x (measured by emulation); range is not an issue for LayerNorm-bounded
// activations, conversions saturate.
#include <cuda.h>
#include <cuda_fp16.h>

#include <map>
#include <vector>

#include "kernels.h"
#include "tc_gemm.h"

namespace cone {

namespace {

constexpr int BM = 128;          // UMMA M
constexpr int BK = 64;           // fp16 elements per stage row = 128 bytes = one swizzle atom row
constexpr int UMMA_K = 16;
constexpr int STAGES = 4;
constexpr int A_BYTES = BM * BK * 2;  // 16 KB
constexpr int TC_THREADS = 192;

// ---------------------------------------------------------------------------------------------- PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "LAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra LAB_DONE;\n\t"
        "bra LAB_WAIT;\n\t"
        "LAB_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]^T
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major operand tile in shared memory, 128-byte swizzle: rows of 128 bytes, 8-row atoms 1024 bytes apart
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);  // start address, 16-byte units
    d |= (uint64_t)1 << 16;                   // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;         // stride byte offset: next 8-row atom
    d |= (uint64_t)1 << 46;                   // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                   // SWIZZLE_128B
    return d;
}
// 32 lanes x 32 columns of fp32 accumulators: thread i of the warp gets TMEM lane (base + i), 32 consecutive columns
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct TcEpilogue {
    const float* bias;   // [N] or null
    const float* R;      // fp32 residual [M, ldr] or null
    int64_t ldr;
    float* C32;          // fp32 output or null
    int64_t ldc32;
    __half* C16;         // fp16 output or null
    int64_t ldc16;
    int relu;
    const float* ln_g;   // fused LayerNorm over the N = BN columns of the row (null = off)
    const float* ln_b;
    float ln_eps;
};

// ------------------------------------------------------------------------------------------ kernel
template <int BN>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, TcEpilogue ep,
               int64_t M, int N, int K) {
    constexpr int B_BYTES = BN * BK * 2;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * A_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(sB + STAGES * B_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tfull = empty + STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t m_tiles = (M + BM - 1) / BM;
    const int n_tiles = N / BN;
    const int64_t tiles = m_tiles * n_tiles;
    const int num_k = K / BK;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull[i], 1);
            mbar_init(&tempty[i], 128);
        }
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
            for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
                const int m0 = (int)(t / n_tiles) * BM, n0 = (int)(t % n_tiles) * BN;
                for (int kb = 0; kb < num_k; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    mbar_expect_tx(&full[stage], A_BYTES + B_BYTES);
                    tma_load_2d(sA + stage * A_BYTES, &tmA, &full[stage], kb * BK, m0);
                    tma_load_2d(sB + stage * B_BYTES, &tmB, &full[stage], kb * BK, n0);
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {  // ------------------------------------------------------------------ MMA issuer
            // instruction descriptor: D fp32, A/B fp16 K-major, N = BN, M = 128
            constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
                mbar_wait(&tempty[acc], acc_phase ^ 1);  // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
                for (int kb = 0; kb < num_k; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(sA + stage * A_BYTES);
                    const uint32_t b_addr = smem_u32(sB + stage * B_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        const uint64_t da = smem_desc_sw128(a_addr + k * UMMA_K * 2);
                        const uint64_t db = smem_desc_sw128(b_addr + k * UMMA_K * 2);
                        umma_f16(d_tmem, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(&empty[stage]);  // frees the smem stage once these MMAs have read it
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit(&tfull[acc]);  // accumulator complete -> epilogue
                if (++acc == 2) {
                    acc = 0;
                    acc_phase ^= 1;
                }
            }
        }
    } else {  // ------------------------------------------------------------------------------- epilogue
        const int quarter = warp & 3;  // TMEM lane quarter this warp may read
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
            const int64_t m0 = (t / n_tiles) * BM;
            const int n0 = (int)(t % n_tiles) * BN;
            mbar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
            const int64_t row = m0 + quarter * 32 + lane;
            const bool row_ok = row < M;
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN);
            float mean = 0.f, rstd = 1.f;
            if (ep.ln_g != nullptr) {  // fused LayerNorm: statistics over the full row (N == BN)
                float s1 = 0.f, s2 = 0.f;
                for (int c = 0; c < BN; c += 32) {
                    float v[32];
                    tmem_ld_32x32(taddr + c, v);
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        float x = v[j];
                        if (ep.bias) x += __ldg(ep.bias + n0 + c + j);
                        if (ep.R && row_ok) x += ep.R[row * ep.ldr + n0 + c + j];
                        s1 += x;
                        s2 = fmaf(x, x, s2);
                    }
                }
                mean = s1 * (1.f / BN);
                const float var = fmaxf(s2 * (1.f / BN) - mean * mean, 0.f);
                rstd = rsqrtf(var + ep.ln_eps);
            }
            for (int c = 0; c < BN; c += 32) {
                float v[32];
                tmem_ld_32x32(taddr + c, v);
                if (row_ok) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        float x = v[j];
                        if (ep.bias) x += __ldg(ep.bias + n0 + c + j);
                        if (ep.R) x += ep.R[row * ep.ldr + n0 + c + j];
                        if (ep.relu) x = fmaxf(x, 0.f);
                        if (ep.ln_g) x = (x - mean) * rstd * __ldg(ep.ln_g + n0 + c + j) + __ldg(ep.ln_b + n0 + c + j);
                        v[j] = x;
                    }
                    if (ep.C32) {
                        float4* o = reinterpret_cast<float4*>(ep.C32 + row * ep.ldc32 + n0 + c);
#pragma unroll
                        for (int j = 0; j < 8; ++j) o[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    }
                    if (ep.C16) {
                        uint4* o = reinterpret_cast<uint4*>(ep.C16 + row * ep.ldc16 + n0 + c);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            __half2 h0 = __floats2half2_rn(v[8 * j], v[8 * j + 1]);
                            __half2 h1 = __floats2half2_rn(v[8 * j + 2], v[8 * j + 3]);
                            __half2 h2 = __floats2half2_rn(v[8 * j + 4], v[8 * j + 5]);
                            __half2 h3 = __floats2half2_rn(v[8 * j + 6], v[8 * j + 7]);
                            uint4 u;
                            u.x = *reinterpret_cast<uint32_t*>(&h0);
                            u.y = *reinterpret_cast<uint32_t*>(&h1);
                            u.z = *reinterpret_cast<uint32_t*>(&h2);
                            u.w = *reinterpret_cast<uint32_t*>(&h3);
                            o[j] = u;
                        }
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

// 2-D fp16 tensor [rows, cols] with row pitch ld (elements); box = [BK cols, box_rows rows], 128-byte swizzle
int make_map(CUtensorMap* map, const __half* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return CONE_ERR_CUDA;
    }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) for [%lld x %lld] ld %lld", (int)r, (long long)rows,
                  (long long)cols, (long long)ld);
        return CONE_ERR_CUDA;
    }
    return CONE_OK;
}

template <int BN>
constexpr size_t tc_smem_bytes() {
    return 1024 + (size_t)STAGES * (A_BYTES + BN * BK * 2) + 16 * sizeof(uint64_t);
}

}  // namespace

struct TcWeights {
    struct W16 {
        __half* ptr;
        CUtensorMap map;
        int N, K, BN;
    };
    std::map<const float*, W16> cache;  // fp16 copies of nn.Linear weights, keyed by the fp32 device pointer
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

static int get_w16(TcWeights* t, const float* W, int N, int K, cudaStream_t s, const TcWeights::W16** out) {
    auto it = t->cache.find(W);
    if (it == t->cache.end() || it->second.N != N || it->second.K != K) {
        TcWeights::W16 w{};
        w.N = N;
        w.K = K;
        w.BN = (N % 256 == 0) ? 256 : 128;
        CONE_CUDA(cudaMalloc(&w.ptr, (size_t)N * K * 2));
        CONE_TRY(f32_to_f16(W, K, w.ptr, N, K, s));
        CONE_TRY(make_map(&w.map, w.ptr, N, K, K, w.BN));
        if (it != t->cache.end()) cudaFree(it->second.ptr);
        t->cache[W] = w;
        it = t->cache.find(W);
    }
    *out = &it->second;
    return CONE_OK;
}

static int tc_gemm_f16_impl(TcWeights* t, const __half* A16, int64_t lda, int64_t M, const float* W, const float* b, int N, int K,
                float* C32, int64_t ldc32, __half* C16, int64_t ldc16, int relu, const float* R, int64_t ldr,
                const float* ln_g, const float* ln_b, cudaStream_t s) {
    CONE_REQUIRE(t != nullptr, "tc_gemm: tensor-core weights not initialised");
    CONE_REQUIRE(tc_gemm_supported(M, N, K), "tc_gemm: unsupported shape M=%lld N=%d K=%d", (long long)M, N, K);
    CONE_REQUIRE((lda % 8) == 0 && (reinterpret_cast<uintptr_t>(A16) & 15) == 0, "tc_gemm: A must be 16-byte aligned");
    CONE_REQUIRE(C32 == nullptr || ((ldc32 % 4) == 0 && (reinterpret_cast<uintptr_t>(C32) & 15) == 0), "tc_gemm: C32 alignment");
    CONE_REQUIRE(C16 == nullptr || ((ldc16 % 8) == 0 && (reinterpret_cast<uintptr_t>(C16) & 15) == 0), "tc_gemm: C16 alignment");
    const TcWeights::W16* w = nullptr;
    CONE_TRY(get_w16(t, W, N, K, s, &w));
    CONE_REQUIRE(ln_g == nullptr || N == w->BN, "tc_gemm: fused LayerNorm needs the whole row in one tile (N=%d)", N);
    CUtensorMap mapA;
    CONE_TRY(make_map(&mapA, A16, M, K, lda, BM));
    TcEpilogue ep{b, R, ldr, C32, ldc32, C16, ldc16, relu, ln_g, ln_b, 1e-5f};
    const int64_t tiles = cdiv64(M, BM) * (N / w->BN);
    const unsigned grid = (unsigned)(tiles < t->num_sms ? tiles : t->num_sms);
    ProfScope ps(s, P_GEMM_TC, 2.0 * (double)M * N * K,
                 2.0 * ((double)M * K + (double)N * K) + (C32 ? 4.0 : 0.0) * M * N + (C16 ? 2.0 : 0.0) * M * N +
                     (R ? 4.0 : 0.0) * M * N);
    if (w->BN == 256) {
        static bool attr = false;
        if (!attr) {
            CONE_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)tc_smem_bytes<256>()));
            attr = true;
        }
        tc_gemm_kernel<256><<<grid, TC_THREADS, tc_smem_bytes<256>(), s>>>(mapA, w->map, ep, M, N, K);
    } else {
        static bool attr = false;
        if (!attr) {
            CONE_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)tc_smem_bytes<128>()));
            attr = true;
        }
        tc_gemm_kernel<128><<<grid, TC_THREADS, tc_smem_bytes<128>(), s>>>(mapA, w->map, ep, M, N, K);
    }
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
    return tc_gemm_f16_impl(t, a16, K, M, W, b, N, K, y, ldy, nullptr, 0, relu, R, ldr, nullptr, nullptr, s);
}

int tc_gemm_f16(TcWeights* t, const uint16_t* A16, int64_t lda, int64_t M, const float* W, const float* b, int N, int K,
                float* C32, int64_t ldc32, uint16_t* C16, int64_t ldc16, int relu, const float* R, int64_t ldr,
                const float* ln_g, const float* ln_b, cudaStream_t s) {
    return tc_gemm_f16_impl(t, reinterpret_cast<const __half*>(A16), lda, M, W, b, N, K, C32, ldc32,
                            reinterpret_cast<__half*>(C16), ldc16, relu, R, ldr, ln_g, ln_b, s);
}

}  // namespace cone

// Stage 3 of the reference, one CTA per query (cone/inference.py:70-91, 103-127, 169-217;
// utils/temporal_nms.py:6-74).  Everything after the fp32 span arithmetic is fp64, as in the reference's
// Python floats; orderings are stable with the insertion ordinal (window rank, then within-window score
// rank) as the secondary key.
#include "kernels.h"

namespace cone {

namespace {

constexpr int NMS_THREADS = 128;
constexpr int MAX_NQ = 8;

// float(f"{x:.4f}") for a value that came from fp32: x*1e4 is exact in fp64 (24+10 significant bits),
// rint is round-half-even like Python's correctly rounded formatting, and n/1e4 is the correctly rounded
// quotient, i.e. the double nearest to the decimal string.
__device__ __forceinline__ double round4(float x) { return rint((double)x * 1e4) / 1e4; }

// compute_temporal_iou (temporal_nms.py:6-22): intersection over the hull, 0 when the hull is empty
__device__ __forceinline__ double hull_iou(double s1, double e1, double s2, double e2) {
    const double inter = fmax(0.0, fmin(e1, e2) - fmax(s1, s2));
    const double hull = fmax(e1, e2) - fmin(s1, s2);
    return hull == 0.0 ? 0.0 : inter / hull;
}

// Greedy NMS over candidates order[0..P) (already sorted, best first).  Writes kept candidate ids.
// Equivalent to temporal_nms.py:45-71 including its "append the last survivor" tail.
__device__ int greedy_nms(const double* st, const double* ed, const int* order, int P, double thd, int max_after,
                          unsigned char* dead, int* kept, int* sh_head) {
    for (int i = threadIdx.x; i < P; i += blockDim.x) dead[i] = 0;
    __syncthreads();
    if (P == 1) {  // `if len(predictions) == 1: return predictions`
        if (threadIdx.x == 0) kept[0] = order[0];
        __syncthreads();
        return max_after >= 0 ? 1 : 0;
    }
    int nkept = 0;
    int head = 0;
    while (nkept < max_after) {
        if (threadIdx.x == 0) {
            int h = head;
            while (h < P && dead[h]) ++h;
            *sh_head = h;
        }
        __syncthreads();
        head = *sh_head;
        if (head >= P) break;
        const int c = order[head];
        const double s1 = st[c], e1 = ed[c];
        for (int t = head + 1 + threadIdx.x; t < P; t += blockDim.x) {
            if (!dead[t]) {
                const int o = order[t];
                if (hull_iou(s1, e1, st[o], ed[o]) > thd) dead[t] = 1;
            }
        }
        if (threadIdx.x == 0) kept[nkept] = c;
        ++nkept;
        ++head;
        __syncthreads();
    }
    __syncthreads();
    return nkept;
}

// stable descending rank by value: position of i = #{j : v[j] > v[i] or (v[j] == v[i] and j < i)}.
// NaN (undefined order in the reference's Python sort) is ranked last so that `order` is always a permutation.
__device__ void stable_rank_desc(const double* v, int M, int* order) {
    for (int i = threadIdx.x; i < M; i += blockDim.x) {
        const double vi = v[i];
        const bool ni = vi != vi;
        int r = 0;
        for (int j = 0; j < M; ++j) {
            const double vj = v[j];
            const bool nj = vj != vj;
            r += (ni || nj) ? ((ni && !nj) || (ni && nj && j < i)) : ((vj > vi) || (vj == vi && j < i));
        }
        order[r] = i;
    }
    __syncthreads();
}

struct Smem {
    double *st, *ed, *sc, *mt, *fu;      // N rows in insertion order
    double *ust, *ued, *usc, *umt, *ufu;  // M unique (st,ed) entries: first position, last value
    int *last, *order, *kept;
    unsigned char* dead;
};

__device__ Smem carve(unsigned char* base, int N) {
    Smem s;
    double* d = reinterpret_cast<double*>(base);
    s.st = d; s.ed = d + N; s.sc = d + 2 * N; s.mt = d + 3 * N; s.fu = d + 4 * N;
    s.ust = d + 5 * N; s.ued = d + 6 * N; s.usc = d + 7 * N; s.umt = d + 8 * N; s.ufu = d + 9 * N;
    int* i = reinterpret_cast<int*>(d + 10 * N);
    s.last = i; s.order = i + N; s.kept = i + 2 * N;
    s.dead = reinterpret_cast<unsigned char*>(i + 3 * N);
    return s;
}

__host__ __device__ inline size_t smem_bytes(int N) { return (size_t)N * (10 * 8 + 3 * 4 + 1) + 64; }

__global__ void __launch_bounds__(NMS_THREADS)
fuse_nms_kernel(const float* __restrict__ pred_spans, const float* __restrict__ prob_fg, const float* __restrict__ match,
                const int32_t* __restrict__ win_start, const int32_t* __restrict__ win_len, int topk, int nq,
                float clip_length, double nms_thd, int max_before, int max_after, int fixed_duration, int sort_windows,
                double* __restrict__ out, int32_t* __restrict__ out_count, double* __restrict__ rows_out,
                int32_t* __restrict__ rows_count) {
    extern __shared__ __align__(16) unsigned char raw[];
    __shared__ int sh_head, sh_M, sh_cnt;
    __shared__ double sh_mm[4];
    const int q = blockIdx.x;
    const int Nmax = topk * nq;
    Smem s = carve(raw, Nmax);

    // present windows form a prefix of the top-k list (absent ones have len 0)
    if (threadIdx.x == 0) {
        int c = 0;
        while (c < topk && win_len[(int64_t)q * topk + c] > 0) ++c;
        sh_cnt = c;
    }
    __syncthreads();
    const int cnt = sh_cnt;
    const int N = cnt * nq;

    // A10: per window, seconds in fp32 (no FMA contraction), stable sort by fp32 score, 4-decimal rounding
    for (int j = threadIdx.x; j < cnt; j += blockDim.x) {
        const int64_t g = (int64_t)q * topk + j;
        // eval_epoch scales by the window's own length (inference.py:75-78), the demo by max_v_l (cone_localizator.py:188)
        const float dur = fixed_duration > 0 ? (float)fixed_duration : (float)win_len[g], vs = (float)win_start[g];
        int idx[MAX_NQ];
        float key[MAX_NQ];
        for (int a = 0; a < nq; ++a) {
            const float k = prob_fg[g * nq + a];
            int pos = a;
            // the demo keeps slot order (cone_localizator.py:186-197): no per-window sort
            while (sort_windows && pos > 0 && key[pos - 1] < k) {  // strict: equal scores keep slot order
                key[pos] = key[pos - 1];
                idx[pos] = idx[pos - 1];
                --pos;
            }
            key[pos] = k;
            idx[pos] = a;
        }
        for (int a = 0; a < nq; ++a) {
            const int64_t p = g * nq + idx[a];
            const float cx = pred_spans[p * 2], w = pred_spans[p * 2 + 1];
            const float hw = __fmul_rn(0.5f, w);
            const float x1 = __fsub_rn(cx, hw), x2 = __fadd_rn(cx, hw);
            const float st = __fmul_rn(__fadd_rn(__fmul_rn(x1, dur), vs), clip_length);
            const float ed = __fmul_rn(__fadd_rn(__fmul_rn(x2, dur), vs), clip_length);
            const int r = j * nq + a;
            s.st[r] = round4(st);
            s.ed[r] = round4(ed);
            s.sc[r] = round4(prob_fg[p]);
            s.mt[r] = round4(match[p]);
        }
    }
    __syncthreads();
    if (rows_out) {
        for (int r = threadIdx.x; r < N; r += blockDim.x) {
            double* o = rows_out + ((int64_t)q * Nmax + r) * 4;
            o[0] = s.st[r]; o[1] = s.ed[r]; o[2] = s.sc[r]; o[3] = s.mt[r];
        }
        if (threadIdx.x == 0 && rows_count) rows_count[q] = N;
    }
    if (N == 0) {
        if (threadIdx.x < 3) out_count[q * 3 + threadIdx.x] = 0;
        return;
    }

    // A11: min-max normalisation of both scores (identity when constant) and their sum
    if (threadIdx.x < 4) {
        const double* v = (threadIdx.x < 2) ? s.sc : s.mt;
        const bool want_max = threadIdx.x & 1;
        double m = v[0];
        for (int r = 1; r < N; ++r) m = want_max ? (v[r] > m ? v[r] : m) : (v[r] < m ? v[r] : m);
        sh_mm[threadIdx.x] = m;
    }
    __syncthreads();
    {
        const double smin = sh_mm[0], smax = sh_mm[1], mmin = sh_mm[2], mmax = sh_mm[3];
        for (int r = threadIdx.x; r < N; r += blockDim.x) {
            const double a = (smin == smax) ? s.sc[r] : (s.sc[r] - smin) / (smax - smin);
            const double b = (mmin == mmax) ? s.mt[r] : (s.mt[r] - mmin) / (mmax - mmin);
            s.fu[r] = a + b;
        }
    }
    __syncthreads();
    // dict keyed by (st, ed): first position, last value
    for (int r = threadIdx.x; r < N; r += blockDim.x) {
        const double a = s.st[r], b = s.ed[r];
        int first = r, last = r;
        for (int t = 0; t < N; ++t) {
            if (s.st[t] == a && s.ed[t] == b) {
                if (t < first) first = t;
                if (t > last) last = t;
            }
        }
        s.last[r] = (first == r) ? last : -1;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int m = 0;
        for (int r = 0; r < N; ++r) {
            const int l = s.last[r];
            if (l >= 0) {
                s.ust[m] = s.st[r]; s.ued[m] = s.ed[r];
                s.usc[m] = s.sc[l]; s.umt[m] = s.mt[l]; s.ufu[m] = s.fu[l];
                ++m;
            }
        }
        sh_M = m;
    }
    __syncthreads();
    const int M = sh_M;

    // A12/A13: three rankings — 0 fusion (idx 2), 1 proposal (idx 0), 2 matching (idx 1)
    for (int mode = 0; mode < 3; ++mode) {
        const double* val = mode == 0 ? s.ufu : (mode == 1 ? s.usc : s.umt);
        stable_rank_desc(val, M, s.order);
        int nk;
        if (nms_thd != -1.0) {
            const int P = M < max_before ? M : max_before;
            nk = greedy_nms(s.ust, s.ued, s.order, P, nms_thd, max_after, s.dead, s.kept, &sh_head);
        } else {
            nk = M < max_after ? M : max_after;
            for (int i = threadIdx.x; i < nk; i += blockDim.x) s.kept[i] = s.order[i];
            __syncthreads();
        }
        for (int i = threadIdx.x; i < nk; i += blockDim.x) {
            const int c = s.kept[i];
            double* o = out + (((int64_t)q * 3 + mode) * max_after + i) * 5;
            o[0] = s.ust[c]; o[1] = s.ued[c]; o[2] = s.usc[c]; o[3] = s.umt[c]; o[4] = s.ufu[c];
        }
        if (threadIdx.x == 0) out_count[q * 3 + mode] = nk;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(NMS_THREADS)
temporal_nms_single_kernel(const double* __restrict__ st, const double* __restrict__ ed,
                           const double* __restrict__ score, int n, double thd, int max_after,
                           int32_t* __restrict__ keep_out, int32_t* __restrict__ n_keep_out) {
    extern __shared__ __align__(16) unsigned char raw[];
    __shared__ int sh_head;
    int* order = reinterpret_cast<int*>(raw);
    int* kept = order + n;
    unsigned char* dead = reinterpret_cast<unsigned char*>(kept + n);
    if (n == 0) {
        if (threadIdx.x == 0) n_keep_out[0] = 0;
        return;
    }
    stable_rank_desc(score, n, order);
    const int nk = greedy_nms(st, ed, order, n, thd, max_after, dead, kept, &sh_head);
    for (int i = threadIdx.x; i < nk; i += blockDim.x) keep_out[i] = kept[i];
    if (threadIdx.x == 0) n_keep_out[0] = nk;
}

}  // namespace

int fuse_nms(const float* pred_spans, const float* prob_fg, const float* match, const int32_t* win_start,
             const int32_t* win_len, int n_queries, int topk, int nq, float clip_length, double nms_thd,
             int max_before_nms, int max_after_nms, double* out, int32_t* out_count, double* rows_out,
             int32_t* rows_count, cudaStream_t s, int fixed_duration, int sort_windows) {
    if (n_queries == 0) return CONE_OK;
    CONE_REQUIRE(nq >= 1 && nq <= MAX_NQ, "fuse_nms: 1..%d moment slots supported", MAX_NQ);
    CONE_REQUIRE(max_after_nms >= 1 && max_before_nms >= 1, "fuse_nms: max_before/after_nms must be >= 1");
    const size_t smem = smem_bytes(topk * nq);
    CONE_REQUIRE(smem <= 200 * 1024, "fuse_nms: %d candidates per query exceed shared memory", topk * nq);
    if (smem > 48 * 1024)
        CONE_CUDA(cudaFuncSetAttribute(fuse_nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ProfScope ps(s, P_NMS);
    fuse_nms_kernel<<<n_queries, NMS_THREADS, smem, s>>>(pred_spans, prob_fg, match, win_start, win_len, topk, nq,
                                                         clip_length, nms_thd, max_before_nms, max_after_nms,
                                                         fixed_duration, sort_windows, out, out_count, rows_out,
                                                         rows_count);
    CONE_LAUNCH_CHECK("fuse_nms");
    return CONE_OK;
}

int temporal_nms_single(const double* st, const double* ed, const double* score, int n, double nms_thd,
                        int max_after_nms, int32_t* keep_out, int32_t* n_keep_out, cudaStream_t s) {
    CONE_REQUIRE(n >= 0 && max_after_nms >= 0, "temporal_nms: negative sizes");
    const size_t smem = (size_t)n * 9 + 64;
    CONE_REQUIRE(smem <= 200 * 1024, "temporal_nms: %d predictions exceed shared memory", n);
    if (smem > 48 * 1024)
        CONE_CUDA(cudaFuncSetAttribute(temporal_nms_single_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    temporal_nms_single_kernel<<<1, NMS_THREADS, smem, s>>>(st, ed, score, n, nms_thd, max_after_nms, keep_out,
                                                           n_keep_out);
    CONE_LAUNCH_CHECK("temporal_nms_single");
    return CONE_OK;
}

}  // namespace cone

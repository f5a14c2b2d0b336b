// Metric kernels: R@K / IoU and window pre-filtering recall straight from the stage-3 output, so that the last
// host loop of eval_epoch (cone/inference.py:333-474) needs only a handful of counters from the device.
//   flavour 0 = standalone_eval/evaluate_mad.py:32-37, 60-104     (fp32 hull IoU, fp32 thresholds, strict >)
//   flavour 1 = standalone_eval/evaluate_ego4d_nlq.py:41-62, 65-117 (fp64 IoU, fp64 thresholds, top-1 IoU for mIoU)
//   window recall = standalone_eval/evaluate_pre_filtered_window.py:30-72
#include "kernels.h"

namespace cone {

namespace {

constexpr int kMaxK = CONE_EVAL_MAX_TOPK;
constexpr int kMaxThr = CONE_EVAL_MAX_THRESHOLDS;

struct EvalSpec {
    int n_topk, n_thr;
    int topk[kMaxK];
    double thr[kMaxThr];
};

// one thread per (query, ranking): walks the <= max_after predictions once, keeping for every threshold the rank of
// the first hit; hit counters are integer atomics (order-independent, so the result is deterministic)
__global__ void recall_kernel(const double* __restrict__ nms, const int32_t* __restrict__ cnt, const double* __restrict__ gt,
                              int n_queries, int max_after, EvalSpec spec, int flavour,
                              unsigned long long* __restrict__ hits, double* __restrict__ top1) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_queries * 3) return;
    const int q = t / 3, m = t - 3 * q;
    const double* rows = nms + ((size_t)q * 3 + m) * max_after * 5;
    int max_k = 0;
    for (int i = 0; i < spec.n_topk; ++i) max_k = spec.topk[i] > max_k ? spec.topk[i] : max_k;
    int n = cnt[q * 3 + m];
    n = n < max_after ? n : max_after;
    n = n < max_k ? n : max_k;  // predicted_times[:max_recall]
    int first_hit[kMaxThr];
    for (int j = 0; j < spec.n_thr; ++j) first_hit[j] = 0x7fffffff;
    const double g0 = gt[2 * q], g1 = gt[2 * q + 1];
    double iou0 = 0.0;
    for (int i = 0; i < n; ++i) {
        const double st = rows[i * 5], ed = rows[i * 5 + 1];
        if (flavour == 0) {
            // torch.tensor(list of Python floats) is float32; _iou works on .float() copies
            const float s = (float)st, e = (float)ed, a = (float)g0, b = (float)g1;
            const float inter = fminf(e, b) - fmaxf(s, a);
            const float hull = fmaxf(e, b) - fminf(s, a);
            const float iou = __fdiv_rn(fmaxf(inter, 0.f), hull);
            for (int j = 0; j < spec.n_thr; ++j)
                if (iou > (float)spec.thr[j] && i < first_hit[j]) first_hit[j] = i;
            if (i == 0) iou0 = (double)iou;
        } else {
            const double inter = fmax(0.0, fmin(ed, g1) - fmax(st, g0));
            const double hull = fmax(0.0, fmax(ed, g1) - fmin(st, g0));
            const double iou = __ddiv_rn(inter, hull);
            for (int j = 0; j < spec.n_thr; ++j)
                if (iou > spec.thr[j] && i < first_hit[j]) first_hit[j] = i;
            if (i == 0) iou0 = iou;
        }
    }
    for (int i = 0; i < spec.n_topk; ++i)
        for (int j = 0; j < spec.n_thr; ++j)
            if (first_hit[j] < spec.topk[i]) atomicAdd(&hits[((size_t)m * spec.n_topk + i) * spec.n_thr + j], 1ull);
    if (top1) top1[q * 3 + m] = iou0;
}

// one thread per query: ground-truth window ids are range(floor(start / sws), ceil(end / sws) + 1) with start / end in
// frames (fp64 division by clip_length, as the Python floats of the reference)
__global__ void window_recall_kernel(const int32_t* __restrict__ ranklist, int stride, const double* __restrict__ gt,
                                     int n_queries, double clip_length, int sws, EvalSpec spec,
                                     unsigned long long* __restrict__ hits) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_queries) return;
    int max_k = 0;
    for (int i = 0; i < spec.n_topk; ++i) max_k = spec.topk[i] > max_k ? spec.topk[i] : max_k;
    const double start = __ddiv_rn(gt[2 * q], clip_length), end = __ddiv_rn(gt[2 * q + 1], clip_length);
    const double lo = floor(__ddiv_rn(start, (double)sws)), hi = ceil(__ddiv_rn(end, (double)sws));  // inclusive
    int first = 0x7fffffff;
    const int n = stride < max_k ? stride : max_k;
    for (int i = 0; i < n; ++i) {
        const int w = ranklist[(size_t)q * stride + i];
        if (w < 0) break;  // -1 padding: the video has fewer windows
        if ((double)w >= lo && (double)w <= hi) {
            first = i;
            break;
        }
    }
    for (int i = 0; i < spec.n_topk; ++i)
        if (first < spec.topk[i]) atomicAdd(&hits[i], 1ull);
}

int make_spec(const int32_t* topk_host, int n_topk, const double* thr_host, int n_thr, EvalSpec& s) {
    CONE_REQUIRE(n_topk >= 1 && n_topk <= kMaxK, "eval: 1..%d recall ranks", kMaxK);
    CONE_REQUIRE(n_thr >= 0 && n_thr <= kMaxThr, "eval: at most %d IoU thresholds", kMaxThr);
    s.n_topk = n_topk;
    s.n_thr = n_thr;
    for (int i = 0; i < n_topk; ++i) {
        CONE_REQUIRE(topk_host[i] >= 1, "eval: recall ranks must be >= 1");
        s.topk[i] = topk_host[i];
    }
    for (int j = 0; j < n_thr; ++j) s.thr[j] = thr_host[j];
    return CONE_OK;
}

}  // namespace

int eval_recall(const double* nms, const int32_t* nms_count, const double* gt, int n_queries, int max_after,
                const int32_t* topk_host, int n_topk, const double* thr_host, int n_thr, int flavour, int64_t* hits,
                double* top1_iou, cudaStream_t s) {
    EvalSpec spec;
    CONE_TRY(make_spec(topk_host, n_topk, thr_host, n_thr, spec));
    CONE_REQUIRE(n_thr >= 1, "eval: at least one IoU threshold");
    CONE_REQUIRE(flavour == 0 || flavour == 1, "eval: flavour 0 (MAD) or 1 (Ego4D)");
    if (n_queries == 0) return CONE_OK;
    const int threads = 128, total = n_queries * 3;
    ProfScope ps(s, P_NMS);
    recall_kernel<<<cdiv(total, threads), threads, 0, s>>>(nms, nms_count, gt, n_queries, max_after, spec, flavour,
                                                           (unsigned long long*)hits, top1_iou);
    CONE_LAUNCH_CHECK("recall_kernel");
    return CONE_OK;
}

int eval_window_recall(const int32_t* ranklist, int ranklist_stride, const double* gt, int n_queries, double clip_length,
                       int max_v_l, const int32_t* topk_host, int n_topk, int64_t* hits, cudaStream_t s) {
    EvalSpec spec;
    CONE_TRY(make_spec(topk_host, n_topk, nullptr, 0, spec));
    CONE_REQUIRE(max_v_l >= 2 && clip_length > 0.0, "eval: bad window length / clip length");
    if (n_queries == 0) return CONE_OK;
    const int threads = 128;
    ProfScope ps(s, P_NMS);
    window_recall_kernel<<<cdiv(n_queries, threads), threads, 0, s>>>(ranklist, ranklist_stride, gt, n_queries, clip_length,
                                                                      max_v_l / 2, spec, (unsigned long long*)hits);
    CONE_LAUNCH_CHECK("window_recall_kernel");
    return CONE_OK;
}

}  // namespace cone

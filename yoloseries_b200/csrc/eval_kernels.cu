// eval_kernels.cu -- the consumers right after the kept rows (SURVEY.md 8f "next" rows):
//   * mAP_v2.compute_tp and its pairwise IoU        utils/mAP.py:18-42, 70-100   (k_compute_tp, k_map_iou)
//   * weighted-box-fusion                           utils/weighted_fusion_bbox.py:41-96, trainer/eval_yolov5.py:44-92
//                                                   (k_wbf)
// Small, latency-bound problems (<= max_det predictions x a few dozen ground-truth boxes per image; a few thousand boxes
// per image for the fusion): one CTA per image / per (image, label) segment, everything else is parallel over images.
#include "ysb_internal.cuh"

#include <climits>

namespace ysb {

// ---- utils/mAP.py:18-42 ------------------------------------------------------------------------------------------------
// numpy, in the arrays' own precision T: areas = w*h, inter = max(0, .) * max(0, .), iou = inter / clip(a1 + a2 - inter,
// 1e-6, 1e7).  a = ground truth row (x1, y1, x2, y2, ...), b = prediction row.
template <typename T>
__device__ __forceinline__ T map_iou(const T *a, const T *b)
{
    const T a1 = (a[2] - a[0]) * (a[3] - a[1]);
    const T a2 = (b[2] - b[0]) * (b[3] - b[1]);
    const T xmin = a[0] > b[0] ? a[0] : b[0], ymin = a[1] > b[1] ? a[1] : b[1];   // np.maximum / np.minimum
    const T xmax = a[2] < b[2] ? a[2] : b[2], ymax = a[3] < b[3] ? a[3] : b[3];
    T w = xmax - xmin, h = ymax - ymin;
    w = w > T(0) ? w : T(0);
    h = h > T(0) ? h : T(0);
    const T inter = w * h;
    T den = (a1 + a2) - inter;
    den = den < T(1e-6) ? T(1e-6) : (den > T(10000000) ? T(10000000) : den);
    return inter / den;
}

template <typename T>
__global__ void __launch_bounds__(256) k_map_iou(const T *__restrict__ b1, int64_t n, int w1, const T *__restrict__ b2,
                                                 int64_t m, int w2, T *__restrict__ out)
{
    const int64_t i = blockIdx.y;
    const int64_t j = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n || j >= m) return;
    out[i * m + j] = map_iou<T>(b1 + i * w1, b2 + j * w2);
}

struct TpThresholds {
    double v[10];
};

// The best ground-truth box of prediction j under the reference's ordering (utils/mAP.py:88-96): pairs with
// iou >= thr[0] and equal label, sorted by IoU descending; np.unique(pred) keeps the first pair of every prediction.
// Equal IoUs: argsort()[::-1] of a stable sort visits the LATER pair first, i.e. the larger ground-truth index.
template <typename T>
__device__ __forceinline__ int best_gt(const T *gt, int n, const T *p, double thr0, T &best_iou)
{
    int bi = -1;
    T bv = T(0);
    for (int i = 0; i < n; ++i) {
        const T *g = gt + static_cast<int64_t>(i) * 5;
        const T v = map_iou<T>(g, p);
        if (static_cast<double>(v) >= thr0 && g[4] == p[5] && (bi < 0 || v >= bv)) {
            bi = i;
            bv = v;
        }
    }
    best_iou = bv;
    return bi;
}

// One CTA per image.  gt rows (x1, y1, x2, y2, cls), pred rows (x1, y1, x2, y2, score, cls); images are concatenated,
// gt_off / pred_off (batch + 1) are row offsets.  first_pred: one int per ground-truth row (workspace).
//   phase 1: every prediction finds its best ground-truth box and bids for it with atomicMin(pred index): the second
//            np.unique (over the ground-truth column of the pred-ordered list) keeps the LOWEST prediction index of
//            every ground-truth box -- not the highest IoU; that is the reference's behaviour and it is kept.
//   phase 2: the winning prediction of every ground-truth box gets tp[j, t] = iou >= thr[t] (float64 comparison).
template <typename T>
__global__ void __launch_bounds__(256) k_compute_tp(const T *__restrict__ gt, const int64_t *__restrict__ gt_off,
                                                    const T *__restrict__ pred, const int64_t *__restrict__ pred_off,
                                                    const __grid_constant__ TpThresholds thr, int32_t *__restrict__ first_pred,
                                                    uint8_t *__restrict__ tp)
{
    const int img = blockIdx.x;
    const int64_t g0 = gt_off[img], p0 = pred_off[img];
    const int n = static_cast<int>(gt_off[img + 1] - g0), m = static_cast<int>(pred_off[img + 1] - p0);
    for (int i = threadIdx.x; i < n; i += blockDim.x) first_pred[g0 + i] = INT_MAX;
    for (int e = threadIdx.x; e < m * 10; e += blockDim.x) tp[p0 * 10 + e] = 0;
    __syncthreads();
    const T *g = gt + g0 * 5;
    for (int j = threadIdx.x; j < m; j += blockDim.x) {
        T v;
        const int bi = best_gt<T>(g, n, pred + (p0 + j) * 6, thr.v[0], v);
        if (bi >= 0) atomicMin(first_pred + g0 + bi, j);
    }
    __syncthreads();
    for (int j = threadIdx.x; j < m; j += blockDim.x) {
        T v;
        const int bi = best_gt<T>(g, n, pred + (p0 + j) * 6, thr.v[0], v);
        if (bi >= 0 && first_pred[g0 + bi] == j) {
#pragma unroll
            for (int t = 0; t < 10; ++t) tp[(p0 + j) * 10 + t] = static_cast<double>(v) >= thr.v[t] ? 1 : 0;
        }
    }
}

cudaError_t launch_map_iou(const void *b1, int64_t n, int w1, const void *b2, int64_t m, int w2, int f64, void *out,
                           cudaStream_t stream)
{
    if (n == 0 || m == 0) return cudaSuccess;
    for (int64_t r0 = 0; r0 < n; r0 += 65535) {
        const int64_t rows = (n - r0) < 65535 ? (n - r0) : 65535;
        const dim3 grid(static_cast<unsigned>((m + 255) / 256), static_cast<unsigned>(rows));
        if (f64)
            k_map_iou<double><<<grid, 256, 0, stream>>>(static_cast<const double *>(b1) + r0 * w1, rows, w1,
                                                        static_cast<const double *>(b2), m, w2,
                                                        static_cast<double *>(out) + r0 * m);
        else
            k_map_iou<float><<<grid, 256, 0, stream>>>(static_cast<const float *>(b1) + r0 * w1, rows, w1,
                                                       static_cast<const float *>(b2), m, w2,
                                                       static_cast<float *>(out) + r0 * m);
    }
    return cudaGetLastError();
}

cudaError_t launch_compute_tp(const void *gt, const int64_t *gt_off, const void *pred, const int64_t *pred_off, int batch,
                              int f64, const double *thr10, int32_t *first_pred, uint8_t *tp, cudaStream_t stream)
{
    if (batch == 0) return cudaSuccess;
    TpThresholds t;
    for (int i = 0; i < 10; ++i) t.v[i] = thr10[i];
    if (f64)
        k_compute_tp<double><<<batch, 256, 0, stream>>>(static_cast<const double *>(gt), gt_off,
                                                        static_cast<const double *>(pred), pred_off, t, first_pred, tp);
    else
        k_compute_tp<float><<<batch, 256, 0, stream>>>(static_cast<const float *>(gt), gt_off,
                                                       static_cast<const float *>(pred), pred_off, t, first_pred, tp);
    return cudaGetLastError();
}

}  // namespace ysb

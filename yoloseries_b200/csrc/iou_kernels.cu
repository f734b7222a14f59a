// iou_kernels.cu -- the IoU routines of utils/bbox_tools.py as standalone kernels (forward only; callers that
// need autograd keep using the torch expressions, see yoloseries_b200/utils/bbox_tools.py).
//   pairwise  NUMBA_F64MIX  numba_iou  utils/bbox_tools.py:12-35   -> (n, m) float64
//   pairwise  F32           gpu_iou    utils/bbox_tools.py:164-190 -> (n, m) float32
//   rowwise   GIoU/DIoU/CIoU           utils/bbox_tools.py:193-339 -> (n) float32, b1 broadcast when it has one row
#include "ysb_internal.cuh"

namespace ysb {

template <bool F64>
__global__ void __launch_bounds__(256) k_pairwise_iou(const float4 *__restrict__ b1, int64_t n, const float4 *__restrict__ b2,
                                                      int64_t m, void *__restrict__ out)
{
    // one output row per blockIdx.y, 256 consecutive columns per block: coalesced stores, b1 row broadcast
    const int64_t i = blockIdx.y;
    const int64_t j = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n || j >= m) return;
    const float4 a = __ldg(b1 + i), b = __ldg(b2 + j);
    if (F64) {
        const OffBox oa = make_offbox(a, 0.0f), ob = make_offbox(b, 0.0f);
        static_cast<double *>(out)[i * m + j] = iou_numba_exact(oa, ob);
    } else {
        static_cast<float *>(out)[i * m + j] = iou_f32(a, b);
    }
}

__global__ void __launch_bounds__(256) k_elementwise_iou(const float4 *__restrict__ b1, int64_t n1, const float4 *__restrict__ b2,
                                                         int64_t n2, int kind, float *__restrict__ out)
{
    const int64_t j = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (j >= n2) return;
    out[j] = iou_kind_f32(kind, __ldg(b1 + (n1 == 1 ? 0 : j)), __ldg(b2 + j));
}

// ---- soft-NMS, utils/nms.py:68-140: one CTA, one pick per iteration (block arg-max + decay sweep) -----------------
__global__ void __launch_bounds__(1024, 1) k_soft_nms(const float4 *__restrict__ boxes, const float *__restrict__ scores_in,
                                                      int m, float thr, int kind, int mode, float sigma, long long cap,
                                                      float *__restrict__ score, float *__restrict__ processed)
{
    __shared__ float s_val[32];
    __shared__ int s_idx[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int j = tid; j < m; j += 1024) { score[j] = scores_in[j]; processed[j] = 0.0f; }
    __syncthreads();
    for (long long it = 0; it < cap; ++it) {
        float bs = -INFINITY;
        int bj = 0x7fffffff;
        for (int j = tid; j < m; j += 1024) {
            const float s = score[j];
            if (s > bs) { bs = s; bj = j; }  // ascending j per thread: first maximum wins
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const float os = __shfl_xor_sync(0xffffffffu, bs, d);
            const int oj = __shfl_xor_sync(0xffffffffu, bj, d);
            if (os > bs || (os == bs && oj < bj)) { bs = os; bj = oj; }
        }
        if (lane == 0) { s_val[warp] = bs; s_idx[warp] = bj; }
        __syncthreads();
        bs = s_val[lane];
        bj = s_idx[lane];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const float os = __shfl_xor_sync(0xffffffffu, bs, d);
            const int oj = __shfl_xor_sync(0xffffffffu, bj, d);
            if (os > bs || (os == bs && oj < bj)) { bs = os; bj = oj; }
        }
        if (!(bs > 0.0f)) break;  // "while score.sum() > 0" for non-negative scores
        if (tid == 0) processed[bj] = bs;
        const float4 b1 = __ldg(boxes + bj);
        for (int j = tid; j < m; j += 1024) {
            const float v = iou_kind_f32(kind, b1, __ldg(boxes + j));
            if (v > thr) {
                const float f = mode == 0 ? __fsub_rn(1.0f, v) : expf(__fdiv_rn(-__fmul_rn(v, v), sigma));
                score[j] = __fmul_rn(score[j], f);
            }
        }
        __syncthreads();
    }
}

__global__ void k_undo_letterbox(float *__restrict__ dets, const int32_t *__restrict__ cnt, int max_det,
                                 const float *__restrict__ info)
{
    const int img = blockIdx.x;
    const int n = cnt[img];
    const float scale = info[img * 5 + 0], pad_top = info[img * 5 + 1], pad_left = info[img * 5 + 2];
    const float hi_y = __fsub_rn(info[img * 5 + 3], 1.0f), hi_x = __fsub_rn(info[img * 5 + 4], 1.0f);
    for (int r = threadIdx.x; r < n; r += blockDim.x) {
        float *row = dets + (static_cast<size_t>(img) * max_det + r) * 6;
        row[0] = fminf(fmaxf(__fdiv_rn(__fsub_rn(row[0], pad_left), scale), 1.0f), hi_x);
        row[2] = fminf(fmaxf(__fdiv_rn(__fsub_rn(row[2], pad_left), scale), 1.0f), hi_x);
        row[1] = fminf(fmaxf(__fdiv_rn(__fsub_rn(row[1], pad_top), scale), 1.0f), hi_y);
        row[3] = fminf(fmaxf(__fdiv_rn(__fsub_rn(row[3], pad_top), scale), 1.0f), hi_y);
    }
}

cudaError_t launch_soft_nms(const float *boxes, const float *scores, int64_t m, float thr, int kind, int mode, float sigma,
                            void *ws, float *processed, cudaStream_t stream)
{
    if (m == 0) return cudaSuccess;
    // a picked box decays itself by exp(-1/sigma) per pick in exponential mode: ~104*sigma picks until float32 underflow
    long long per_box = 64;
    if (mode == 1 && sigma > 0.0f) {
        const double need = 104.0 * static_cast<double>(sigma) + 2.0;
        if (need > 64.0) per_box = need < 1.0e6 ? static_cast<long long>(need) : 1000000ll;
    }
    const long long cap = per_box * m + 1024;
    k_soft_nms<<<1, 1024, 0, stream>>>(reinterpret_cast<const float4 *>(boxes), scores, static_cast<int>(m), thr, kind, mode,
                                       sigma, cap, static_cast<float *>(ws), processed);
    return cudaGetLastError();
}

cudaError_t launch_undo_letterbox(float *dets, const int32_t *cnt, int batch, int max_det, const float *info, cudaStream_t stream)
{
    if (batch == 0) return cudaSuccess;
    k_undo_letterbox<<<batch, 128, 0, stream>>>(dets, cnt, max_det, info);
    return cudaGetLastError();
}

cudaError_t launch_pairwise_iou(const float *b1, int64_t n, const float *b2, int64_t m, int kind, void *out, cudaStream_t stream)
{
    if (n == 0 || m == 0) return cudaSuccess;
    for (int64_t r0 = 0; r0 < n; r0 += 65535) {  // gridDim.y limit
        const int64_t rows = (n - r0) < 65535 ? (n - r0) : 65535;
        const dim3 grid(static_cast<unsigned>((m + 255) / 256), static_cast<unsigned>(rows));
        const float4 *a = reinterpret_cast<const float4 *>(b1) + r0;
        if (kind == YSB_IOU_NUMBA_F64MIX)
            k_pairwise_iou<true><<<grid, 256, 0, stream>>>(a, rows, reinterpret_cast<const float4 *>(b2), m,
                                                           static_cast<double *>(out) + r0 * m);
        else
            k_pairwise_iou<false><<<grid, 256, 0, stream>>>(a, rows, reinterpret_cast<const float4 *>(b2), m,
                                                            static_cast<float *>(out) + r0 * m);
    }
    return cudaGetLastError();
}

cudaError_t launch_elementwise_iou(const float *b1, int64_t n1, const float *b2, int64_t n2, int kind, float *out,
                                   cudaStream_t stream)
{
    if (n2 == 0) return cudaSuccess;
    k_elementwise_iou<<<static_cast<unsigned>((n2 + 255) / 256), 256, 0, stream>>>(
        reinterpret_cast<const float4 *>(b1), n1, reinterpret_cast<const float4 *>(b2), n2, kind, out);
    return cudaGetLastError();
}

}  // namespace ysb

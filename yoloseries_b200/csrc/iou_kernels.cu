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

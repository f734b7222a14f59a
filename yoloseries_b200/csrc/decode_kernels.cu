// decode_kernels.cu -- the do_inference decode as ONE kernel per family: raw heads -> (b, N, C') rows.
//
// Replaces the ~12 ATen launches and three full-tensor copies per stage of trainer/eval_yolov5.py:188-209
// (and eval_yolov7.py:131-151, eval_yolox.py:135-150, eval_yolov8.py:76-102, eval_retinanet.py:59-75,
// eval_fcos.py:126-161).  This is the API-compatibility path (XEvaluator.do_inference returns the full tensor);
// the fused path (ysb_postprocess) never materialises these rows.
//
// A CTA owns 32 consecutive candidates of one image; sigmoid/exp values are produced into a shared-memory tile
// [32][C'] and written out with fully coalesced stores (32 rows of one image are contiguous in the output).
// Planes heads are read coalesced along positions (lane = candidate), rows heads along the row (lane = column).
#include "ysb_internal.cuh"

namespace ysb {

constexpr int kDecRows = 32;
constexpr int kDecThreads = 256;

// class / objectness column value of the decoded row for raw heads
__device__ __forceinline__ float decoded_class_value(const Plan &P, int img, int cand, int l, int k)
{
    const LevelDesc &lv = P.lv[l];
    const int r = cand - lv.cand_off;
    if (P.layout == LAYOUT_PLANES) {
        const int a = r / lv.hw, pos = r - a * lv.hw;
        return sigmoid_ref(__ldg(lv.p0 + (static_cast<size_t>(img * P.A + a) * P.cls_nch + P.cls_ch + k) * lv.hw + pos));
    }
    return sigmoid_ref(__ldg(lv.p0 + (static_cast<size_t>(img) * lv.img_rows + r) * P.row_w_in + P.cls_col_in + k));
}

__device__ __forceinline__ float decoded_obj_value(const Plan &P, int img, int cand, int l)
{
    const LevelDesc &lv = P.lv[l];
    const int r = cand - lv.cand_off;
    if (P.obj_src == 1)
        return sigmoid_ref(__ldg(P.lv[0].p1 + (static_cast<size_t>(img) * P.N + cand) * P.reg_row_w + 4));
    if (P.layout == LAYOUT_PLANES) {
        const int a = r / lv.hw, pos = r - a * lv.hw;
        const float *src = (P.obj_src == 2 ? lv.p2 : lv.p0);
        return sigmoid_ref(__ldg(src + (static_cast<size_t>(img * P.A + a) * P.obj_nch + P.obj_ch) * lv.hw + pos));
    }
    return sigmoid_ref(__ldg(lv.p0 + (static_cast<size_t>(img) * lv.img_rows + r) * P.row_w_in + P.obj_col_in));
}

__global__ void __launch_bounds__(kDecThreads) k_decode_rows(const __grid_constant__ Plan P, float *__restrict__ out)
{
    extern __shared__ float tile[];  // [kDecRows][row_w]
    const int img = blockIdx.y;
    const int c0 = blockIdx.x * kDecRows;
    const int nrows = min(kDecRows, P.N - c0);
    const int rw = P.row_w;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int kWarps = kDecThreads / 32;
    if (P.layout == LAYOUT_PLANES) {
        // lane = candidate, warps stride over the class columns: each load instruction covers 32 neighbouring
        // positions of one plane (candidates of a CTA may straddle an anchor/level boundary; still correct)
        const int cand = c0 + lane;
        if (lane < nrows) {
            const int l = find_level(P, cand);
            for (int k = warp; k < P.C; k += kWarps) tile[lane * rw + P.cls_col + k] = decoded_class_value(P, img, cand, l, k);
            if (warp == 0) {
                const float4 b = decode_box_cols(P, img, cand);
                float *t = tile + lane * rw + P.box_col;
                t[0] = b.x; t[1] = b.y; t[2] = b.z; t[3] = b.w;
            }
            if (warp == 1 % kWarps && P.obj_col >= 0) tile[lane * rw + P.obj_col] = decoded_obj_value(P, img, cand, l);
        }
    } else {
        // warp = candidate row, lanes stride over its columns (contiguous in the channels-last head)
        for (int r = warp; r < nrows; r += kWarps) {
            const int cand = c0 + r;
            const int l = find_level(P, cand);
            for (int k = lane; k < P.C; k += 32) tile[r * rw + P.cls_col + k] = decoded_class_value(P, img, cand, l, k);
            if (lane == 0) {
                const float4 b = decode_box_cols(P, img, cand);
                float *t = tile + r * rw + P.box_col;
                t[0] = b.x; t[1] = b.y; t[2] = b.z; t[3] = b.w;
            }
            if (lane == 1 && P.obj_col >= 0) tile[r * rw + P.obj_col] = decoded_obj_value(P, img, cand, l);
        }
    }
    __syncthreads();
    float *dst = out + (static_cast<size_t>(img) * P.N + c0) * rw;
    const int nfl = nrows * rw;
    for (int e = threadIdx.x; e < nfl; e += kDecThreads) dst[e] = tile[e];
}

cudaError_t launch_decode(const Plan &P, float *d_out, cudaStream_t stream)
{
    if (P.batch == 0 || P.N == 0) return cudaSuccess;
    const size_t smem = static_cast<size_t>(kDecRows) * P.row_w * sizeof(float);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_decode_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return e;
    }
    const dim3 grid((P.N + kDecRows - 1) / kDecRows, P.batch);
    k_decode_rows<<<grid, kDecThreads, smem, stream>>>(P, d_out);
    return cudaGetLastError();
}

}  // namespace ysb

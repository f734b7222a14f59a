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

// Address of class logit 0 of a candidate and the stride between consecutive classes (hoisted out of the class loop:
// the level lookup and the two integer divisions are per candidate, not per element).
struct CandAddr {
    const float *cls;
    size_t cstride;
    const float *obj;  // objectness / centerness / conf logit, nullptr when the family has none
};

__device__ __forceinline__ CandAddr cand_addr(const Plan &P, int img, int cand)
{
    CandAddr c;
    const int l = find_level(P, cand);
    const LevelDesc &lv = P.lv[l];
    const int r = cand - lv.cand_off;
    c.obj = nullptr;
    if (P.layout == LAYOUT_PLANES) {
        const int a = r / lv.hw, pos = r - a * lv.hw;
        c.cls = lv.p0 + (static_cast<size_t>(img * P.A + a) * P.cls_nch + P.cls_ch) * lv.hw + pos;
        c.cstride = static_cast<size_t>(lv.hw);
        if (P.obj_col >= 0)
            c.obj = (P.obj_src == 2 ? lv.p2 : lv.p0) + (static_cast<size_t>(img * P.A + a) * P.obj_nch + P.obj_ch) * lv.hw + pos;
    } else {
        const float *row = lv.p0 + (static_cast<size_t>(img) * lv.img_rows + r) * P.row_w_in;
        c.cls = row + P.cls_col_in;
        c.cstride = 1;
        if (P.obj_col >= 0)
            c.obj = P.obj_src == 1 ? P.lv[0].p1 + (static_cast<size_t>(img) * P.N + cand) * P.reg_row_w + 4 : row + P.obj_col_in;
    }
    return c;
}

// ROWS = candidates per CTA: 32 for channels-last heads (their rows are contiguous anyway), 128 for NCHW planes so that
// every plane access of a CTA covers 512 contiguous bytes.
template <int ROWS>
__global__ void __launch_bounds__(kDecThreads) k_decode_rows(const __grid_constant__ Plan P, float *__restrict__ out,
                                                              int64_t rows_total, int64_t row_offset)
{
    extern __shared__ __align__(16) float tile[];  // [ROWS][row_w]
    const int img = blockIdx.y;
    const int c0 = blockIdx.x * ROWS;
    const int nrows = min(ROWS, P.N - c0);
    const int rw = P.row_w;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (P.layout == LAYOUT_PLANES) {
        // thread = (candidate, class phase): consecutive threads take consecutive positions of one plane; four
        // independent loads in flight per thread, then four sigmoids
        constexpr int kPhases = kDecThreads / ROWS;  // 2 for ROWS = 128
        const int rr = threadIdx.x % ROWS, ph = threadIdx.x / ROWS;
        if (rr < nrows) {
            const CandAddr ca = cand_addr(P, img, c0 + rr);
            float *tc = tile + rr * rw + P.cls_col;
            const size_t step = static_cast<size_t>(kPhases) * ca.cstride;
            int k = ph;
            const float *src = ca.cls + static_cast<size_t>(k) * ca.cstride;
            for (; k + 3 * kPhases < P.C; k += 4 * kPhases, src += 4 * step) {
                const float v0 = __ldg(src), v1 = __ldg(src + step), v2 = __ldg(src + 2 * step), v3 = __ldg(src + 3 * step);
                tc[k] = sigmoid_ref(v0);
                tc[k + kPhases] = sigmoid_ref(v1);
                tc[k + 2 * kPhases] = sigmoid_ref(v2);
                tc[k + 3 * kPhases] = sigmoid_ref(v3);
            }
            for (; k < P.C; k += kPhases, src += step) tc[k] = sigmoid_ref(__ldg(src));
            if (ph == 0 && ca.obj) tile[rr * rw + P.obj_col] = sigmoid_ref(__ldg(ca.obj));
        }
        if (P.family == YSB_YOLOV8) {
            // DFL boxes: four threads per candidate (one side each), 64 candidates per round
            for (int base = 0; base < nrows; base += kDecThreads / 4) {
                const int r4 = base + (threadIdx.x >> 2), side = threadIdx.x & 3;
                float sv = 0.0f;
                if (r4 < nrows) sv = v8_side_value(P, img, c0 + r4, side);
                const unsigned qb = lane & ~3u;
                const float s0 = __shfl_sync(0xffffffffu, sv, qb), s1 = __shfl_sync(0xffffffffu, sv, qb + 1);
                const float s2 = __shfl_sync(0xffffffffu, sv, qb + 2), s3 = __shfl_sync(0xffffffffu, sv, qb + 3);
                if (side == 0 && r4 < nrows) {
                    const float4 b = tta_undo(P, v8_box_from_sides(P, c0 + r4, s0, s1, s2, s3));
                    float *t = tile + r4 * rw + P.box_col;
                    t[0] = b.x; t[1] = b.y; t[2] = b.z; t[3] = b.w;
                }
            }
        } else if (threadIdx.x >= kDecThreads - ROWS && threadIdx.x - (kDecThreads - ROWS) < nrows) {
            const int rb = threadIdx.x - (kDecThreads - ROWS);
            const float4 b = tta_undo(P, decode_box_cols(P, img, c0 + rb));
            float *t = tile + rb * rw + P.box_col;
            t[0] = b.x; t[1] = b.y; t[2] = b.z; t[3] = b.w;
        }
    } else {
        // eight threads per candidate row (32 rows x 8): a warp reads 4 rows x 32 contiguous bytes per step, every
        // fetched sector fully used; four loads in flight per thread.  Lane-per-candidate box / objectness decode.
        const int r = threadIdx.x >> 3, q = threadIdx.x & 7;
        if (r < nrows) {
            const CandAddr ca = cand_addr(P, img, c0 + r);
            float *tc = tile + r * rw + P.cls_col;
            int k = q;
            for (; k + 24 < P.C; k += 32) {
                const float v0 = __ldg(ca.cls + k), v1 = __ldg(ca.cls + k + 8), v2 = __ldg(ca.cls + k + 16), v3 = __ldg(ca.cls + k + 24);
                tc[k] = sigmoid_ref(v0);
                tc[k + 8] = sigmoid_ref(v1);
                tc[k + 16] = sigmoid_ref(v2);
                tc[k + 24] = sigmoid_ref(v3);
            }
            for (; k < P.C; k += 8) tc[k] = sigmoid_ref(__ldg(ca.cls + k));
        }
        if (warp == 0 && lane < nrows) {
            const int cand = c0 + lane;
            const float4 b = tta_undo(P, decode_box_cols(P, img, cand));
            float *t = tile + lane * rw + P.box_col;
            t[0] = b.x; t[1] = b.y; t[2] = b.z; t[3] = b.w;
            if (P.obj_col >= 0) {
                const CandAddr ca = cand_addr(P, img, cand);
                t[P.obj_col - P.box_col] = sigmoid_ref(__ldg(ca.obj));
            }
        }
    }
    __syncthreads();
    float *dst = out + (static_cast<size_t>(img) * rows_total + row_offset + c0) * rw;
    const int nfl = nrows * rw;
    if (((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(tile)) & 15u) == 0) {
        const int nv = nfl >> 2;
        for (int e = threadIdx.x; e < nv; e += kDecThreads)
            reinterpret_cast<float4 *>(dst)[e] = reinterpret_cast<const float4 *>(tile)[e];
        for (int e = (nv << 2) + threadIdx.x; e < nfl; e += kDecThreads) dst[e] = tile[e];
    } else {
        for (int e = threadIdx.x; e < nfl; e += kDecThreads) dst[e] = tile[e];
    }
}

template <int ROWS>
static cudaError_t launch_decode_t(const Plan &P, float *d_out, int64_t rows_total, int64_t row_offset, cudaStream_t stream)
{
    const size_t smem = static_cast<size_t>(ROWS) * P.row_w * sizeof(float);
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_decode_rows<ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return e;
    }
    const dim3 grid((P.N + ROWS - 1) / ROWS, P.batch);
    k_decode_rows<ROWS><<<grid, kDecThreads, smem, stream>>>(P, d_out, rows_total, row_offset);
    return cudaGetLastError();
}

// rows_total / row_offset: the pass fills rows [row_offset, row_offset + N) of a (batch, rows_total, C') tensor
cudaError_t launch_decode(const Plan &P, float *d_out, int64_t rows_total, int64_t row_offset, cudaStream_t stream)
{
    if (P.batch == 0 || P.N == 0) return cudaSuccess;
    if (P.layout == LAYOUT_PLANES && static_cast<size_t>(128) * P.row_w * sizeof(float) <= 100 * 1024)
        return launch_decode_t<128>(P, d_out, rows_total, row_offset, stream);
    return launch_decode_t<32>(P, d_out, rows_total, row_offset, stream);
}

}  // namespace ysb

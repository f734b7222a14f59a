// iou_kernels.cu -- the IoU routines of utils/bbox_tools.py as standalone kernels, forward and (for the flavours the
// reference's losses differentiate through) backward; see yoloseries_b200/utils/bbox_tools.py.
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

// ---- backward of the row-wise GIoU / DIoU / CIoU (the loss-side callers: loss/yolov5_loss.py:110, yolov7_loss.py:130,
// yolov8_loss.py:306, loss.py:109 differentiate through gpu_CIoU) -----------------------------------------------------
// Reverse-mode walk of the torch graph of utils/bbox_tools.py:193-339 with torch's sub-gradient conventions:
// maximum/minimum split the gradient evenly at ties, clamp passes it where the input is inside [min, max] (bounds
// included), abs uses sign(x) (0 at 0); alpha of CIoU is a constant (computed under no_grad, :334-335).
struct Grad4 {
    float x1, y1, x2, y2;
};
__device__ __forceinline__ void max_bwd(float p, float q, float g, float &gp, float &gq)
{
    const float wp = p > q ? 1.0f : (p == q ? 0.5f : 0.0f);
    gp += g * wp;
    gq += g * (1.0f - wp);
}
__device__ __forceinline__ void min_bwd(float p, float q, float g, float &gp, float &gq)
{
    const float wp = p < q ? 1.0f : (p == q ? 0.5f : 0.0f);
    gp += g * wp;
    gq += g * (1.0f - wp);
}

__device__ __forceinline__ void iou_kind_backward(int kind, float4 a, float4 b, float G, Grad4 &ga, Grad4 &gb)
{
    const float eps_u = kind == YSB_CIOU ? 1e-9f : 1e-6f;  // clamp of the union (:308 / :216, :260)
    const float w1 = a.z - a.x, h1 = a.w - a.y, w2 = b.z - b.x, h2 = b.w - b.y;
    const float ix1 = fmaxf(a.x, b.x), iy1 = fmaxf(a.y, b.y), ix2 = fminf(a.z, b.z), iy2 = fminf(a.w, b.w);
    const float tw = ix2 - ix1, th = iy2 - iy1;
    const float iw = fmaxf(tw, 0.0f), ih = fmaxf(th, 0.0f);
    const float inter = iw * ih;
    const float u_raw = w1 * h1 + w2 * h2 - inter;
    const float uc = fmaxf(u_raw, eps_u);
    const float iou = inter / uc;
    const float cx1 = fminf(a.x, b.x), cy1 = fminf(a.y, b.y), cx2 = fmaxf(a.z, b.z), cy2 = fmaxf(a.w, b.w);
    const float cw = cx2 - cx1, ch = cy2 - cy1;

    float d_iou = 0.0f, d_uraw = 0.0f, d_cw = 0.0f, d_ch = 0.0f;
    float d_w1 = 0.0f, d_h1 = 0.0f, d_w2 = 0.0f, d_h2 = 0.0f;
    float d_dx = 0.0f, d_dy = 0.0f;  // centre differences (box1 - box2)
    if (kind == YSB_GIOU) {
        // g = iou - |C - U| / |clamp(C, 1e-6)|
        const float C = cw * ch;
        const float cc = fmaxf(C, 1e-6f);
        const float diff = C - u_raw;
        const float num = fabsf(diff), den = fabsf(cc);
        d_iou = G;
        const float d_num = -G / den, d_den = G * num / (den * den);
        const float sgn = diff > 0.0f ? 1.0f : (diff < 0.0f ? -1.0f : 0.0f);
        const float d_diff = d_num * sgn;
        float d_C = d_diff;
        d_uraw += -d_diff;
        const float d_cc = d_den * (cc > 0.0f ? 1.0f : (cc < 0.0f ? -1.0f : 0.0f));
        if (C >= 1e-6f) d_C += d_cc;
        d_cw += d_C * ch;
        d_ch += d_C * cw;
    } else {
        // DIoU: clamp(iou - rho2 / clamp(c2, 1e-6), -1, 1);  CIoU: iou - (rho2 / clamp(c2, 1e-9) + v * alpha)
        const float eps_c = kind == YSB_CIOU ? 1e-9f : 1e-6f;
        const float c2 = cw * cw + ch * ch;
        const float c2c = fmaxf(c2, eps_c);
        const float dx = (a.z + a.x) * 0.5f - (b.z + b.x) * 0.5f;
        const float dy = (a.w + a.y) * 0.5f - (b.w + b.y) * 0.5f;
        const float rho2 = dx * dx + dy * dy;
        float g_pre = G;
        float v = 0.0f, alpha = 0.0f, delta = 0.0f, r1 = 0.0f, r2 = 0.0f, h1c = 1.0f, h2c = 1.0f;
        if (kind == YSB_DIOU) {
            const float pre = iou - rho2 / c2c;
            if (!(pre >= -1.0f && pre <= 1.0f)) g_pre = 0.0f;
        } else {
            h1c = fmaxf(h1, 1e-9f);
            h2c = fmaxf(h2, 1e-9f);
            r1 = w1 / h1c;
            r2 = w2 / h2c;
            delta = atanf(r1) - atanf(r2);
            v = 0.40528473456935109f * (delta * delta);
            alpha = v / fmaxf(1.0f - iou + v, 1e-9f);
        }
        d_iou = g_pre;
        const float d_pen = -g_pre;
        const float d_rho2 = d_pen / c2c;
        const float d_c2c = -d_pen * rho2 / (c2c * c2c);
        const float d_c2 = c2 >= eps_c ? d_c2c : 0.0f;
        d_cw += 2.0f * cw * d_c2;
        d_ch += 2.0f * ch * d_c2;
        d_dx = 2.0f * dx * d_rho2;
        d_dy = 2.0f * dy * d_rho2;
        if (kind == YSB_CIOU) {
            const float d_v = -g_pre * alpha;
            const float d_delta = d_v * 0.40528473456935109f * 2.0f * delta;
            const float d_r1 = d_delta / (1.0f + r1 * r1), d_r2 = -d_delta / (1.0f + r2 * r2);
            d_w1 += d_r1 / h1c;
            if (h1 >= 1e-9f) d_h1 += -d_r1 * w1 / (h1c * h1c);
            d_w2 += d_r2 / h2c;
            if (h2 >= 1e-9f) d_h2 += -d_r2 * w2 / (h2c * h2c);
        }
    }
    // iou = inter / clamp(u_raw, eps)
    float d_inter = d_iou / uc;
    const float d_uc = -d_iou * inter / (uc * uc);
    if (u_raw >= eps_u) d_uraw += d_uc;
    d_w1 += d_uraw * h1;
    d_h1 += d_uraw * w1;
    d_w2 += d_uraw * h2;
    d_h2 += d_uraw * w2;
    d_inter -= d_uraw;
    const float d_tw = tw >= 0.0f ? d_inter * ih : 0.0f;
    const float d_th = th >= 0.0f ? d_inter * iw : 0.0f;
    // leaves
    ga = Grad4{-d_w1 + 0.5f * d_dx, -d_h1 + 0.5f * d_dy, d_w1 + 0.5f * d_dx, d_h1 + 0.5f * d_dy};
    gb = Grad4{-d_w2 - 0.5f * d_dx, -d_h2 - 0.5f * d_dy, d_w2 - 0.5f * d_dx, d_h2 - 0.5f * d_dy};
    max_bwd(a.x, b.x, -d_tw, ga.x1, gb.x1);  // ix1 = max(ax1, bx1), tw = ix2 - ix1
    min_bwd(a.z, b.z, d_tw, ga.x2, gb.x2);   // ix2 = min(ax2, bx2)
    max_bwd(a.y, b.y, -d_th, ga.y1, gb.y1);
    min_bwd(a.w, b.w, d_th, ga.y2, gb.y2);
    min_bwd(a.x, b.x, -d_cw, ga.x1, gb.x1);  // cx1 = min, cw = cx2 - cx1
    max_bwd(a.z, b.z, d_cw, ga.x2, gb.x2);
    min_bwd(a.y, b.y, -d_ch, ga.y1, gb.y1);
    max_bwd(a.w, b.w, d_ch, ga.y2, gb.y2);
}

__global__ void __launch_bounds__(256) k_elementwise_iou_backward(const float4 *__restrict__ b1, int64_t n1,
                                                                  const float4 *__restrict__ b2, int64_t n2, int kind,
                                                                  const float *__restrict__ grad_out, float4 *__restrict__ g1,
                                                                  float4 *__restrict__ g2)
{
    const int64_t j = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    Grad4 ga{0.f, 0.f, 0.f, 0.f}, gb{0.f, 0.f, 0.f, 0.f};
    if (j < n2) {
        iou_kind_backward(kind, __ldg(b1 + (n1 == 1 ? 0 : j)), __ldg(b2 + j), __ldg(grad_out + j), ga, gb);
        if (g2) g2[j] = make_float4(gb.x1, gb.y1, gb.x2, gb.y2);
    }
    if (!g1) return;
    if (n1 != 1) {
        if (j < n2) g1[j] = make_float4(ga.x1, ga.y1, ga.x2, ga.y2);
        return;
    }
    // broadcast box1: sum over the rows (warp shuffle, then one atomic per warp and component; g1 zeroed by the launcher)
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        ga.x1 += __shfl_xor_sync(0xffffffffu, ga.x1, d);
        ga.y1 += __shfl_xor_sync(0xffffffffu, ga.y1, d);
        ga.x2 += __shfl_xor_sync(0xffffffffu, ga.x2, d);
        ga.y2 += __shfl_xor_sync(0xffffffffu, ga.y2, d);
    }
    if ((threadIdx.x & 31) == 0) {
        float *o = reinterpret_cast<float *>(g1);
        atomicAdd(o + 0, ga.x1);
        atomicAdd(o + 1, ga.y1);
        atomicAdd(o + 2, ga.x2);
        atomicAdd(o + 3, ga.y2);
    }
}

// ---- backward of the pairwise gpu_iou (utils/bbox_tools.py:164-190; differentiated by loss/yolox_loss.py:133 and
// loss/yolov7_loss.py:312, whose label assignment runs with grad enabled: `torch.no_grad()` at yolox_loss.py:92 is a bare
// statement, not a decorator) --------------------------------------------------------------------------------------
// Same conventions as above: min/max split ties evenly, clamp(min=0) / clamp(1e-9) pass the gradient on the bound.
__device__ __forceinline__ void iou_f32_backward(float4 a, float4 b, float G, Grad4 &ga, Grad4 &gb)
{
    const float w1 = a.z - a.x, h1 = a.w - a.y, w2 = b.z - b.x, h2 = b.w - b.y;
    const float tw = fminf(a.z, b.z) - fmaxf(a.x, b.x), th = fminf(a.w, b.w) - fmaxf(a.y, b.y);
    const float iw = fmaxf(tw, 0.0f), ih = fmaxf(th, 0.0f);
    const float inter = iw * ih;
    const float u_raw = w1 * h1 + w2 * h2 - inter;
    const float uc = fmaxf(u_raw, 1e-9f);
    float d_inter = G / uc;
    const float d_uraw = u_raw >= 1e-9f ? -G * inter / (uc * uc) : 0.0f;
    d_inter -= d_uraw;
    const float d_w1 = d_uraw * h1, d_h1 = d_uraw * w1, d_w2 = d_uraw * h2, d_h2 = d_uraw * w2;
    const float d_tw = tw >= 0.0f ? d_inter * ih : 0.0f;
    const float d_th = th >= 0.0f ? d_inter * iw : 0.0f;
    ga = Grad4{-d_w1, -d_h1, d_w1, d_h1};
    gb = Grad4{-d_w2, -d_h2, d_w2, d_h2};
    max_bwd(a.x, b.x, -d_tw, ga.x1, gb.x1);
    min_bwd(a.z, b.z, d_tw, ga.x2, gb.x2);
    max_bwd(a.y, b.y, -d_th, ga.y1, gb.y1);
    min_bwd(a.w, b.w, d_th, ga.y2, gb.y2);
}

// g1[i] = sum_j d(G_ij * iou_ij)/d b1_i: one CTA per row i, threads stride over the columns (coalesced reads of G).
__global__ void __launch_bounds__(256) k_pairwise_iou_bwd_rows(const float4 *__restrict__ b1, const float4 *__restrict__ b2,
                                                               int64_t m, const float *__restrict__ G, float4 *__restrict__ g1)
{
    __shared__ double red[8][4];
    const int64_t i = blockIdx.x;
    const float4 a = __ldg(b1 + i);
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    for (int64_t j = threadIdx.x; j < m; j += 256) {
        Grad4 ga, gb;
        iou_f32_backward(a, __ldg(b2 + j), __ldg(G + i * m + j), ga, gb);
        s0 += ga.x1; s1 += ga.y1; s2 += ga.x2; s3 += ga.y2;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, d);
        s1 += __shfl_xor_sync(0xffffffffu, s1, d);
        s2 += __shfl_xor_sync(0xffffffffu, s2, d);
        s3 += __shfl_xor_sync(0xffffffffu, s3, d);
    }
    if ((threadIdx.x & 31) == 0) {
        double *r = red[threadIdx.x >> 5];
        r[0] = s0; r[1] = s1; r[2] = s2; r[3] = s3;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t[4] = {0.0, 0.0, 0.0, 0.0};
        for (int w = 0; w < 8; ++w)
            for (int c = 0; c < 4; ++c) t[c] += red[w][c];
        g1[i] = make_float4(static_cast<float>(t[0]), static_cast<float>(t[1]), static_cast<float>(t[2]), static_cast<float>(t[3]));
    }
}

// g2[j] = sum_i ...: a CTA owns 32 columns, its 8 thread rows stride over the rows i (each warp reads 32 consecutive
// G values of one row), partial sums meet in shared memory.
__global__ void __launch_bounds__(256) k_pairwise_iou_bwd_cols(const float4 *__restrict__ b1, int64_t n,
                                                               const float4 *__restrict__ b2, int64_t m,
                                                               const float *__restrict__ G, float4 *__restrict__ g2)
{
    __shared__ double red[8][32][4];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t j = static_cast<int64_t>(blockIdx.x) * 32 + tx;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    if (j < m) {
        const float4 b = __ldg(b2 + j);
        for (int64_t i = ty; i < n; i += 8) {
            Grad4 ga, gb;
            iou_f32_backward(__ldg(b1 + i), b, __ldg(G + i * m + j), ga, gb);
            s0 += gb.x1; s1 += gb.y1; s2 += gb.x2; s3 += gb.y2;
        }
    }
    double *r = red[ty][tx];
    r[0] = s0; r[1] = s1; r[2] = s2; r[3] = s3;
    __syncthreads();
    if (ty == 0 && j < m) {
        double t[4] = {0.0, 0.0, 0.0, 0.0};
        for (int w = 0; w < 8; ++w)
            for (int c = 0; c < 4; ++c) t[c] += red[w][tx][c];
        g2[j] = make_float4(static_cast<float>(t[0]), static_cast<float>(t[1]), static_cast<float>(t[2]), static_cast<float>(t[3]));
    }
}

cudaError_t launch_pairwise_iou_backward(const float *b1, int64_t n, const float *b2, int64_t m, const float *grad_out,
                                         float *g1, float *g2, cudaStream_t stream)
{
    if (n == 0 || m == 0) {  // an empty sum: zero gradients for whichever side has rows
        cudaError_t e = cudaSuccess;
        if (g1 && n > 0) e = cudaMemsetAsync(g1, 0, static_cast<size_t>(n) * 16, stream);
        if (e == cudaSuccess && g2 && m > 0) e = cudaMemsetAsync(g2, 0, static_cast<size_t>(m) * 16, stream);
        return e;
    }
    const float4 *a = reinterpret_cast<const float4 *>(b1), *b = reinterpret_cast<const float4 *>(b2);
    if (g1) k_pairwise_iou_bwd_rows<<<static_cast<unsigned>(n), 256, 0, stream>>>(a, b, m, grad_out, reinterpret_cast<float4 *>(g1));
    if (g2)
        k_pairwise_iou_bwd_cols<<<static_cast<unsigned>((m + 31) / 32), 256, 0, stream>>>(a, n, b, m, grad_out,
                                                                                          reinterpret_cast<float4 *>(g2));
    return cudaGetLastError();
}

cudaError_t launch_elementwise_iou_backward(const float *b1, int64_t n1, const float *b2, int64_t n2, int kind,
                                            const float *grad_out, float *g1, float *g2, cudaStream_t stream)
{
    if (n2 == 0) return cudaSuccess;
    if (g1 && n1 == 1) {
        cudaError_t e = cudaMemsetAsync(g1, 0, 4 * sizeof(float), stream);
        if (e != cudaSuccess) return e;
    }
    k_elementwise_iou_backward<<<static_cast<unsigned>((n2 + 255) / 256), 256, 0, stream>>>(
        reinterpret_cast<const float4 *>(b1), n1, reinterpret_cast<const float4 *>(b2), n2, kind, grad_out,
        reinterpret_cast<float4 *>(g1), reinterpret_cast<float4 *>(g2));
    return cudaGetLastError();
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

// filter_kernels.cu -- K1: fused (sigmoid) + filter + class pick + stream compaction.
//
// Replaces the head of every evaluator's numba_nms method (trainer/eval_yolov5.py:265-286 and the same block
// in eval_yolov7/yolox/yolov8/retinanet/retinanet_experiment/fcos; per-family operators in SURVEY.md 8a-2)
// together with the sigmoid part of do_inference for the channels that feed the score.
//
// HBM-bound streaming pass: every class/objectness logit is read exactly once (box channels are NOT read here;
// the few candidates that reach NMS get their boxes decoded in the NMS kernel).  Per candidate the kernel keeps
// the two largest class logits and the first arg-max, evaluates sigmoid only for the winner (sigmoid is
// monotone, float32 rounding is monotone) and falls back to evaluating every near-tied class exactly when the
// runner-up is within a guard band -- so score and class id are bit-identical to "sigmoid all, multiply all,
// max/argmax" as the reference does, at ~5 ALU ops per logit instead of ~25.
//
// Output: one 64-bit sort key per survivor, written with warp-aggregated atomics (order inside an image is
// irrelevant: the key carries the candidate index, see pack_key).
#include "ysb_internal.cuh"

namespace ysb {

__device__ __forceinline__ float4 ldg_stream4(const float *p)
{
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ float ldg_stream1(const float *p)
{
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

__device__ __forceinline__ void top2_update(float c, int k, float &m1, float &m2, int &k0)
{
    m2 = fmaxf(m2, fminf(m1, c));
    k0 = (c > m1) ? k : k0;
    m1 = fmaxf(m1, c);
}

// Decide one candidate.  m1/m2/k0: largest, second largest class value and first index of the largest; objv:
// objectness logit (or probability when DECODED); read_cls(k) re-reads class value k (slow path only).
// Returns true when the candidate survives; pre_pass reports the pre-mask (FCOS top-k needs its count).
template <bool DECODED, class ReadCls>
__device__ __forceinline__ bool decide_candidate(const Plan &P, float m1, float m2, int k0, float objv,
                                                 ReadCls read_cls, float &score, int &cls, bool &pre_pass)
{
    pre_pass = false;
    const float mult = P.use_obj ? (DECODED ? objv : sigmoid_ref(objv)) : 1.0f;
    if (P.pre_kind == PRE_OBJ && !(mult >= P.conf_thr)) return false;
    float smax, p, guard_lo;
    if (DECODED) {
        smax = m1;
        p = P.use_obj ? __fmul_rn(m1, mult) : m1;
        guard_lo = __fmul_rn(m1, 1.0f - 2.0e-6f);  // products of values this close may round together
        if (!(m1 > 0.0f) || !(p > 1.0e-30f)) guard_lo = -INFINITY;
    } else {
        const float e = expf(-m1);
        smax = __fdiv_rn(1.0f, __fadd_rn(1.0f, e));
        p = P.use_obj ? __fmul_rn(smax, mult) : smax;
        // d/dx ln sigmoid(x) = sigmoid(-x) >= sigmoid(-m1) on [m1-d, m1]: a logit gap d = 8e-6 * (1 + e^m1)
        // guarantees a relative sigmoid gap > 4e-6, far above the <= 1e-6 the implementation can blur.
        guard_lo = m1 - 8.0e-6f * (1.0f + __fdiv_rn(1.0f, e));
        if (m1 < -80.0f || !(p > 1.0e-30f) || !(guard_lo == guard_lo)) guard_lo = -INFINITY;
    }
    cls = k0;
    if (!(m2 < guard_lo)) {
        // near-tie (or denormal range): literal evaluation of every class that could reach the maximum
        float best = -INFINITY;
        int kb = 0;
        float sbest = -INFINITY;
        for (int k = 0; k < P.C; ++k) {
            const float c = read_cls(k);
            if (!(c >= guard_lo)) continue;
            const float sk = DECODED ? c : sigmoid_ref(c);
            const float pk = P.use_obj ? __fmul_rn(sk, mult) : sk;
            if (pk > best) { best = pk; kb = k; }
            sbest = fmaxf(sbest, sk);
        }
        p = best;
        cls = kb;
        smax = sbest;
    }
    switch (P.pre_kind) {
    case PRE_OBJ_X_MAX: pre_pass = p >= P.conf_thr; break;  // fl(obj * max cls) == max fl(obj * cls)
    case PRE_MAXCLS: pre_pass = smax >= P.cls_thr; break;
    case PRE_ANY_GT: pre_pass = smax > P.pre_thr; break;
    default: pre_pass = true; break;
    }
    if (!pre_pass) return false;
    score = p;
    return P.post_strict ? (p > P.cls_thr) : (p >= P.cls_thr);
}

// Warp-aggregated append of up to VEC keys per thread (okm: bit j set when k[j] is a survivor).
template <int VEC>
__device__ __forceinline__ void emit_keys(const uint64_t (&k)[VEC], unsigned okm, int npre, uint32_t smax_bits,
                                          uint32_t smin_inv, uint64_t *keys_img, int64_t cap, int32_t *cnt, bool count_pre)
{
    const unsigned lane = threadIdx.x & 31u;
    const int nk = __popc(okm);
    int incl = nk;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= static_cast<unsigned>(d)) incl += v;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (count_pre) {
        const int tp = __reduce_add_sync(0xffffffffu, npre);
        if (lane == 0 && tp) atomicAdd(cnt + 1, tp);
    }
    if (total == 0) return;
    const uint32_t wmax = __reduce_max_sync(0xffffffffu, smax_bits);
    const uint32_t wmin = __reduce_max_sync(0xffffffffu, smin_inv);
    int base = 0;
    if (lane == 0) {
        base = atomicAdd(cnt, total);
        atomicMax(reinterpret_cast<unsigned int *>(cnt + 2), wmax);
        atomicMax(reinterpret_cast<unsigned int *>(cnt + 3), wmin);
    }
    base = __shfl_sync(0xffffffffu, base, 0);
    const int64_t at = static_cast<int64_t>(base) + (incl - nk);
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        const int64_t slot = at + __popc(okm & ((1u << j) - 1u));
        if (((okm >> j) & 1u) && slot < cap) keys_img[slot] = k[j];
    }
}

// -------------------------------------------------------------------------------------------------------
// planes layout (NCHW heads: YOLOv5, YOLOX, YOLOv8, FCOS).  One thread = VEC consecutive positions of one
// (image, anchor); per class plane a warp reads 32*VEC*4 contiguous bytes (512 B with 128-bit loads).
// -------------------------------------------------------------------------------------------------------
template <int VEC>
__global__ void __launch_bounds__(256) k_filter_planes(const __grid_constant__ Plan P, uint64_t *__restrict__ keys,
                                                       int64_t key_cap, int32_t *__restrict__ counts)
{
    const int img = blockIdx.y;
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = u < P.units_per_img;
    uint64_t out[VEC];
    unsigned okm = 0u;
    int npre = 0;
    uint32_t smax_bits = 0u, smin_inv = 0u;
#pragma unroll
    for (int j = 0; j < VEC; ++j) out[j] = 0ull;
    if (active) {
        int l = 0;
#pragma unroll
        for (int i = 1; i < YSB_MAX_LEVELS; ++i)
            if (i < P.L && u >= P.lv[i].unit_off) l = i;
        const LevelDesc &lv = P.lv[l];
        const int upa = lv.hw / VEC;  // units per anchor
        const int ru = u - lv.unit_off;
        const int a = ru / upa;
        const int pos = (ru - a * upa) * VEC;
        const int cand0 = lv.cand_off + a * lv.hw + pos;
        const size_t hw = static_cast<size_t>(lv.hw);
        const float *cls = lv.p0 + (static_cast<size_t>(img * P.A + a) * P.cls_nch + P.cls_ch) * hw + pos;

        float m1[VEC], m2[VEC];
        int k0[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) { m1[j] = -INFINITY; m2[j] = -INFINITY; k0[j] = 0; }

        constexpr int U = 8;  // independent 128-bit loads in flight per thread
        int k = 0;
        for (; k + U <= P.C; k += U) {
            float v[U][VEC];
#pragma unroll
            for (int q = 0; q < U; ++q) {
                if (VEC == 4) {
                    const float4 t = ldg_stream4(cls + static_cast<size_t>(k + q) * hw);
                    v[q][0] = t.x; v[q][1 % VEC] = t.y; v[q][2 % VEC] = t.z; v[q][3 % VEC] = t.w;
                } else {
                    v[q][0] = ldg_stream1(cls + static_cast<size_t>(k + q) * hw);
                }
            }
#pragma unroll
            for (int q = 0; q < U; ++q)
#pragma unroll
                for (int j = 0; j < VEC; ++j) top2_update(v[q][j], k + q, m1[j], m2[j], k0[j]);
        }
        for (; k < P.C; ++k) {
            float v[VEC];
            if (VEC == 4) {
                const float4 t = ldg_stream4(cls + static_cast<size_t>(k) * hw);
                v[0] = t.x; v[1 % VEC] = t.y; v[2 % VEC] = t.z; v[3 % VEC] = t.w;
            } else {
                v[0] = ldg_stream1(cls + static_cast<size_t>(k) * hw);
            }
#pragma unroll
            for (int j = 0; j < VEC; ++j) top2_update(v[j], k, m1[j], m2[j], k0[j]);
        }
        float objv[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) objv[j] = 0.0f;
        if (P.use_obj) {
            const float *ob = (P.obj_src == 2 ? lv.p2 : lv.p0) +
                              (static_cast<size_t>(img * P.A + a) * P.obj_nch + P.obj_ch) * hw + pos;
            if (VEC == 4) {
                const float4 t = ldg_stream4(ob);
                objv[0] = t.x; objv[1 % VEC] = t.y; objv[2 % VEC] = t.z; objv[3 % VEC] = t.w;
            } else {
                objv[0] = ldg_stream1(ob);
            }
        }
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            float score;
            int c;
            bool pre;
            const float *cj = cls + j;
            const bool ok = decide_candidate<false>(
                P, m1[j], m2[j], k0[j], objv[j], [&](int kk) { return __ldg(cj + static_cast<size_t>(kk) * hw); },
                score, c, pre);
            npre += pre ? 1 : 0;
            if (ok) {
                out[j] = pack_key(score, static_cast<uint32_t>(cand0 + j), static_cast<uint32_t>(c));
                okm |= 1u << j;
                const uint32_t sb = __float_as_uint(score);
                smax_bits = max(smax_bits, sb);
                smin_inv = max(smin_inv, ~sb);
            }
        }
    }
    emit_keys<VEC>(out, okm, npre, smax_bits, smin_inv, keys + static_cast<int64_t>(img) * key_cap, key_cap,
                   counts + img * 4, P.pre_kind == PRE_ANY_GT);
}

// -------------------------------------------------------------------------------------------------------
// rows layout (channels-last heads: YOLOv7, RetinaNet cls; and the decoded (b, N, C') tensor of any family).
// A CTA stages a tile of whole rows in shared memory with coalesced 128-bit loads (rows are 340 B / 320 B and
// only 4-byte aligned individually, the tile is contiguous), then one thread scans one row.  The shared row
// stride is forced odd so that the 32 rows a warp scans sit in 32 different banks.
// -------------------------------------------------------------------------------------------------------
constexpr int kRowsTile = 128;

__global__ void __launch_bounds__(kRowsTile) k_filter_rows(const __grid_constant__ Plan P, uint64_t *__restrict__ keys,
                                                           int64_t key_cap, int32_t *__restrict__ counts)
{
    extern __shared__ float tile[];
    const int img = blockIdx.y;
    const int t = blockIdx.x;
    int l = 0;
#pragma unroll
    for (int i = 1; i < YSB_MAX_LEVELS; ++i)
        if (i < P.L && t >= P.lv[i].unit_off) l = i;
    const LevelDesc &lv = P.lv[l];
    const int rows_l = (l + 1 < P.L ? P.lv[l + 1].cand_off : P.N) - lv.cand_off;
    const int r0 = (t - lv.unit_off) * kRowsTile;
    const int nrows = min(kRowsTile, rows_l - r0);
    const int rw = P.row_w_in;
    const int sstride = rw | 1;
    const float *src = lv.p0 + (static_cast<size_t>(img) * lv.img_rows + r0) * rw;
    const int nfl = nrows * rw;
    if ((reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
        const int nv = nfl >> 2;
        for (int i = threadIdx.x; i < nv; i += kRowsTile) {
            const float4 v = ldg_stream4(src + 4 * i);
            int e = 4 * i;
            int row = e / rw, col = e - row * rw;
            const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                tile[row * sstride + col] = vv[j];
                if (++col == rw) { col = 0; ++row; }
            }
        }
        for (int e = (nv << 2) + threadIdx.x; e < nfl; e += kRowsTile) {
            const int row = e / rw, col = e - row * rw;
            tile[row * sstride + col] = ldg_stream1(src + e);
        }
    } else {
        for (int e = threadIdx.x; e < nfl; e += kRowsTile) {
            const int row = e / rw, col = e - row * rw;
            tile[row * sstride + col] = ldg_stream1(src + e);
        }
    }
    __syncthreads();

    uint64_t out[1] = {0ull};
    unsigned okm = 0u;
    int npre = 0;
    uint32_t smax_bits = 0u, smin_inv = 0u;
    if (static_cast<int>(threadIdx.x) < nrows) {
        const float *row = tile + threadIdx.x * sstride;
        const float *cls = row + P.cls_col_in;
        float m1 = -INFINITY, m2 = -INFINITY;
        int k0 = 0;
#pragma unroll 8
        for (int k = 0; k < P.C; ++k) top2_update(cls[k], k, m1, m2, k0);
        const int cand = lv.cand_off + r0 + threadIdx.x;
        float objv = 0.0f;
        if (P.use_obj) {
            if (P.obj_src == 1)  // RetinaNet-exp: conf logit is the last column of the reg tensor
                objv = __ldg(P.lv[0].p1 + (static_cast<size_t>(img) * P.N + cand) * P.reg_row_w + 4);
            else
                objv = row[P.obj_col_in];
        }
        float score;
        int c;
        bool pre, ok;
        if (P.input_kind == YSB_INPUT_DECODED_ROWS)
            ok = decide_candidate<true>(P, m1, m2, k0, objv, [&](int kk) { return cls[kk]; }, score, c, pre);
        else
            ok = decide_candidate<false>(P, m1, m2, k0, objv, [&](int kk) { return cls[kk]; }, score, c, pre);
        npre = pre ? 1 : 0;
        if (ok) {
            out[0] = pack_key(score, static_cast<uint32_t>(cand), static_cast<uint32_t>(c));
            okm = 1u;
            smax_bits = __float_as_uint(score);
            smin_inv = ~smax_bits;
        }
    }
    emit_keys<1>(out, okm, npre, smax_bits, smin_inv, keys + static_cast<int64_t>(img) * key_cap, key_cap,
                 counts + img * 4, P.pre_kind == PRE_ANY_GT);
}

// -------------------------------------------------------------------------------------------------------
// host launchers
// -------------------------------------------------------------------------------------------------------
cudaError_t launch_filter(const Plan &P, int vec, uint64_t *d_keys, int64_t key_cap, int32_t *d_counts,
                          cudaStream_t stream)
{
    cudaError_t e = cudaMemsetAsync(d_counts, 0, sizeof(int32_t) * 4 * static_cast<size_t>(P.batch), stream);
    if (e != cudaSuccess) return e;
    if (P.batch == 0 || P.N == 0) return cudaSuccess;
    if (P.layout == LAYOUT_PLANES) {
        const dim3 grid((P.units_per_img + 255) / 256, P.batch);
        if (vec == 4)
            k_filter_planes<4><<<grid, 256, 0, stream>>>(P, d_keys, key_cap, d_counts);
        else
            k_filter_planes<1><<<grid, 256, 0, stream>>>(P, d_keys, key_cap, d_counts);
    } else {
        const size_t smem = static_cast<size_t>(kRowsTile) * (P.row_w_in | 1) * sizeof(float);
        if (smem > 48 * 1024) {
            e = cudaFuncSetAttribute(k_filter_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
            if (e != cudaSuccess) return e;
        }
        const dim3 grid(P.units_per_img, P.batch);
        k_filter_rows<<<grid, kRowsTile, smem, stream>>>(P, d_keys, key_cap, d_counts);
    }
    return cudaGetLastError();
}

}  // namespace ysb

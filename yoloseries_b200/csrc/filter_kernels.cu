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
//
// Kernels in this file (dispatch: launch_filter at the bottom):
//   k_filter_planes_v4     NCHW heads whose levels are all 128-bit loadable (YOLOv5 / YOLOX / YOLOv8)   -- product path
//   k_filter_planes        NCHW heads with levels that are not (FCOS' 5x5 map); generic                   -- product path
//   k_filter_rows          channels-last heads and the decoded (b, N, C') tensor (YOLOv7, RetinaNet)      -- product path
//   k_filter_multilabel    hyp['mutil_label']: one key per (candidate, class)                              -- product path
//   k_filter_planes_async / _bulk / _tma   cp.async ring, 1-D bulk-copy TMA ring, 2-D tensor-map TMA ring  -- profiling
//       variants selected by YSB_FILTER_VARIANT (same survivor sets, slower on this access pattern; profiles/README.md)
#include <cuda.h>

#include <cstdlib>
#include <cstring>

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
__device__ __forceinline__ float4 ldg_stream4_256(const float *p)
{
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v4.f32 {%0,%1,%2,%3}, [%4];"
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
        // near-tie (or denormal range): literal evaluation of every class that could reach the maximum.  Values are
        // re-read 8 at a time with independent loads (a load-per-iteration loop would serialise 80 DRAM round trips
        // and stall the whole warp for tens of microseconds).
        float best = -INFINITY;
        int kb = 0;
        float sbest = -INFINITY;
        for (int kc = 0; kc < P.C; kc += 8) {
            float cv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) cv[j] = (kc + j < P.C) ? read_cls(kc + j) : -INFINITY;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float c = cv[j];
                if (kc + j < P.C && c >= guard_lo) {
                    const float sk = DECODED ? c : sigmoid_ref(c);
                    const float pk = P.use_obj ? __fmul_rn(sk, mult) : sk;
                    if (pk > best) { best = pk; kb = kc + j; }
                    sbest = fmaxf(sbest, sk);
                }
            }
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
        // the running extrema rarely change after the first warps of an image: look before touching them
        volatile unsigned int *ext = reinterpret_cast<volatile unsigned int *>(cnt + 2);
        if (wmax > ext[0]) atomicMax(reinterpret_cast<unsigned int *>(cnt + 2), wmax);
        if (wmin > ext[1]) atomicMax(reinterpret_cast<unsigned int *>(cnt + 3), wmin);
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
template <int VEC, int U = 8, int THREADS = 256, int MINB = 1, int HINT = 0, int PROBE = 0>
__global__ void __launch_bounds__(THREADS, MINB) k_filter_planes(const __grid_constant__ Plan P, uint64_t *__restrict__ keys,
                                                                 int64_t key_cap, int32_t *__restrict__ counts)
{
    // blockIdx.x = image: CTAs that run at the same time append to DIFFERENT images' counters (the per-image
    // atomicAdd is the only contended address of the kernel)
    const bool img_fast = gridDim.x == static_cast<unsigned>(P.batch) && gridDim.y != static_cast<unsigned>(P.batch);
    const int img = img_fast ? blockIdx.x : blockIdx.y;
    const int u = (img_fast ? blockIdx.y : blockIdx.x) * blockDim.x + threadIdx.x;
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
        const bool vec4 = VEC == 4 && lv.vec == 4;  // levels whose planes are not 16-byte sized use one position per unit
        const int upos = vec4 ? 4 : 1;
        const int upa = lv.hw / upos;  // units per anchor
        const int ru = u - lv.unit_off;
        const int a = ru / upa;
        const int pos = (ru - a * upa) * upos;
        const int cand0 = lv.cand_off + a * lv.hw + pos;
        const size_t hw = static_cast<size_t>(lv.hw);
        const float *cls = lv.p0 + (static_cast<size_t>(img * P.A + a) * P.cls_nch + P.cls_ch) * hw + pos;

        float m1[VEC], m2[VEC];
        int k0[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) { m1[j] = -INFINITY; m2[j] = -INFINITY; k0[j] = 0; }

        // U independent 128-bit loads in flight per thread
        int k = 0;
        for (; k + U <= P.C; k += U) {
            float v[U][VEC];
#pragma unroll
            for (int q = 0; q < U; ++q) {
                if (vec4) {
                    const float4 t = HINT ? ldg_stream4_256(cls + static_cast<size_t>(k + q) * hw) : ldg_stream4(cls + static_cast<size_t>(k + q) * hw);
                    v[q][0] = t.x; v[q][1 % VEC] = t.y; v[q][2 % VEC] = t.z; v[q][3 % VEC] = t.w;
                } else {
                    v[q][0] = ldg_stream1(cls + static_cast<size_t>(k + q) * hw);
#pragma unroll
                    for (int j = 1; j < VEC; ++j) v[q][j] = -INFINITY;
                }
            }
#pragma unroll
            for (int q = 0; q < U; ++q)
#pragma unroll
                for (int j = 0; j < VEC; ++j) {
                    if (PROBE) m1[j] = fmaxf(m1[j], v[q][j]);  // profiling aid: same loads, 1 ALU op per logit
                    else top2_update(v[q][j], k + q, m1[j], m2[j], k0[j]);
                }
        }
        for (; k < P.C; ++k) {
            float v[VEC];
            if (vec4) {
                const float4 t = ldg_stream4(cls + static_cast<size_t>(k) * hw);
                v[0] = t.x; v[1 % VEC] = t.y; v[2 % VEC] = t.z; v[3 % VEC] = t.w;
            } else {
                v[0] = ldg_stream1(cls + static_cast<size_t>(k) * hw);
#pragma unroll
                for (int j = 1; j < VEC; ++j) v[j] = -INFINITY;
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
            if (vec4) {
                const float4 t = ldg_stream4(ob);
                objv[0] = t.x; objv[1 % VEC] = t.y; objv[2 % VEC] = t.z; objv[3 % VEC] = t.w;
            } else {
                objv[0] = ldg_stream1(ob);
            }
        }
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            if (j >= upos) break;
            float score;
            int c;
            bool pre;
            const float *cj = cls + j;
            const bool ok = decide_candidate<false>(
                P, m1[j], m2[j], k0[j], objv[j], [&](int kk) { return __ldg(cj + static_cast<size_t>(kk) * hw); },
                score, c, pre);
            npre += pre ? 1 : 0;
            if (ok) {
                out[j] = pack_key(score, static_cast<uint32_t>(P.cand_base + cand0 + j), static_cast<uint32_t>(c));
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
// planes layout, every level 128-bit loadable (the common case: H*W % 4 == 0 on all levels, 16-byte aligned heads).
// Same decomposition and results as k_filter_planes<4>; specialised so that
//   * there is no per-level scalar fallback path (the generic kernel if-converts both paths and issues both),
//   * plane addresses are 32-bit element offsets from the (image, anchor) block (2 integer ops per load instead of 6),
//   * the last C % U class planes and the objectness plane are ONE batch of independent loads (the generic kernel walks
//     the remainder one dependent load at a time and fetches objectness after the whole class scan).
// -------------------------------------------------------------------------------------------------------
template <int U, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k_filter_planes_v4(const __grid_constant__ Plan P, uint64_t *__restrict__ keys,
                                                                    int64_t key_cap, int32_t *__restrict__ counts)
{
    const int img = blockIdx.x;  // image-fastest grid: concurrent CTAs append to different per-image counters
    const int u = blockIdx.y * THREADS + threadIdx.x;
    const bool active = u < P.units_per_img;
    uint64_t out[4] = {0ull, 0ull, 0ull, 0ull};
    unsigned okm = 0u;
    int npre = 0;
    uint32_t smax_bits = 0u, smin_inv = 0u;
    if (active) {
        int l = 0;
#pragma unroll
        for (int i = 1; i < YSB_MAX_LEVELS; ++i)
            if (i < P.L && u >= P.lv[i].unit_off) l = i;
        const LevelDesc &lv = P.lv[l];
        const uint32_t hw = static_cast<uint32_t>(lv.hw);
        const int upa = lv.hw >> 2;  // units per anchor
        const int ru = u - lv.unit_off;
        const int a = ru / upa;
        const int pos = (ru - a * upa) << 2;
        const int cand0 = lv.cand_off + a * lv.hw + pos;
        const float *cls = lv.p0 + (static_cast<size_t>(img * P.A + a) * P.cls_nch + P.cls_ch) * hw + pos;
        const int C = P.C;

        float m1[4], m2[4], objv[4];
        int k0[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { m1[j] = -INFINITY; m2[j] = -INFINITY; k0[j] = 0; objv[j] = 0.0f; }

        int k = 0;
        uint32_t off = 0;  // k * hw, in elements: an (image, anchor) block is far below 2^32 floats
        for (; k + U <= C; k += U, off += U * hw) {
            float4 v[U];
#pragma unroll
            for (int q = 0; q < U; ++q) v[q] = ldg_stream4_256(cls + (off + q * hw));
#pragma unroll
            for (int q = 0; q < U; ++q) {
                top2_update(v[q].x, k + q, m1[0], m2[0], k0[0]);
                top2_update(v[q].y, k + q, m1[1], m2[1], k0[1]);
                top2_update(v[q].z, k + q, m1[2], m2[2], k0[2]);
                top2_update(v[q].w, k + q, m1[3], m2[3], k0[3]);
            }
        }
        {
            // tail batch: the remaining (< U) class planes and the objectness plane, all in flight together
            const int rem = C - k;
            float4 v[U];
            float4 ov = make_float4(0.f, 0.f, 0.f, 0.f);
            if (P.use_obj) {
                const float *ob = (P.obj_src == 2 ? lv.p2 : lv.p0) +
                                  (static_cast<size_t>(img * P.A + a) * P.obj_nch + P.obj_ch) * hw + pos;
                ov = ldg_stream4_256(ob);
            }
#pragma unroll
            for (int q = 0; q < U - 1; ++q)
                if (q < rem) v[q] = ldg_stream4_256(cls + (off + q * hw));
#pragma unroll
            for (int q = 0; q < U - 1; ++q)
                if (q < rem) {
                    top2_update(v[q].x, k + q, m1[0], m2[0], k0[0]);
                    top2_update(v[q].y, k + q, m1[1], m2[1], k0[1]);
                    top2_update(v[q].z, k + q, m1[2], m2[2], k0[2]);
                    top2_update(v[q].w, k + q, m1[3], m2[3], k0[3]);
                }
            objv[0] = ov.x; objv[1] = ov.y; objv[2] = ov.z; objv[3] = ov.w;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float score;
            int c;
            bool pre;
            const float *cj = cls + j;
            const bool ok = decide_candidate<false>(
                P, m1[j], m2[j], k0[j], objv[j], [&](int kk) { return __ldg(cj + static_cast<size_t>(kk) * hw); }, score, c, pre);
            npre += pre ? 1 : 0;
            if (ok) {
                out[j] = pack_key(score, static_cast<uint32_t>(P.cand_base + cand0 + j), static_cast<uint32_t>(c));
                okm |= 1u << j;
                const uint32_t sb = __float_as_uint(score);
                smax_bits = max(smax_bits, sb);
                smin_inv = max(smin_inv, ~sb);
            }
        }
    }
    emit_keys<4>(out, okm, npre, smax_bits, smin_inv, keys + static_cast<int64_t>(img) * key_cap, key_cap, counts + img * 4,
                 P.pre_kind == PRE_ANY_GT);
}

// -------------------------------------------------------------------------------------------------------
// planes layout, bulk-async version (the default when the heads are 16-byte aligned and H*W % 4 == 0).
//
// Persistent, warp-specialised: one CTA per SM = CW consumer warps + 1 producer warp.  A work item is 128*CW
// consecutive positions of one (image, anchor).  The producer streams the item's class planes (then its
// objectness plane) through a shared-memory ring with cp.async.bulk -- the TMA engine, one row of 512*CW bytes
// per plane, PL planes per ring stage -- signalling mbarriers; each consumer thread owns 4 consecutive positions
// (one 128-bit shared-memory load per plane) and runs the top-2 / arg-max scan.  Bytes in flight no longer depend
// on registers or occupancy: the ring keeps ~200 KB per SM outstanding.
// -------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void *src_gmem, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
                 "l"(src_gmem), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

struct BulkCfg {
    int stages;
    int nq;                // planes streamed per candidate: C class planes, then the objectness plane when use_obj
    int items_per_img;
    int item_off[YSB_MAX_LEVELS];  // first item of each level inside an image
    int blocks_per_anchor[YSB_MAX_LEVELS];
};

struct BulkItem {
    int img, l, a, pos0, npos;
};

template <int kItemPos, typename CFG>
__device__ __forceinline__ BulkItem bulk_item(const Plan &P, const CFG &cfg, int it)
{
    BulkItem w;
    w.img = it / cfg.items_per_img;
    const int r = it - w.img * cfg.items_per_img;
    int l = 0;
#pragma unroll
    for (int i = 1; i < YSB_MAX_LEVELS; ++i)
        if (i < P.L && r >= cfg.item_off[i]) l = i;
    w.l = l;
    const int rr = r - cfg.item_off[l];
    w.a = rr / cfg.blocks_per_anchor[l];
    w.pos0 = (rr - w.a * cfg.blocks_per_anchor[l]) * kItemPos;
    w.npos = min(kItemPos, P.lv[l].hw - w.pos0);
    return w;
}

// global address of streamed plane q at position pos0: q < C is class plane q, q == C the objectness plane
__device__ __forceinline__ const float *bulk_plane(const Plan &P, const BulkItem &w, int q)
{
    const LevelDesc &lv = P.lv[w.l];
    const size_t hw = static_cast<size_t>(lv.hw);
    if (q == P.C)
        return (P.obj_src == 2 ? lv.p2 : lv.p0) + (static_cast<size_t>(w.img * P.A + w.a) * P.obj_nch + P.obj_ch) * hw + w.pos0;
    return lv.p0 + (static_cast<size_t>(w.img * P.A + w.a) * P.cls_nch + P.cls_ch + q) * hw + w.pos0;
}

template <int CW, int PL>
__global__ void __launch_bounds__((CW + 1) * 32, 1)
k_filter_planes_bulk(const __grid_constant__ Plan P, const __grid_constant__ BulkCfg cfg, int total_items,
                     uint64_t *__restrict__ keys, int64_t key_cap, int32_t *__restrict__ counts)
{
    constexpr int kItemPos = 128 * CW;            // positions per item
    constexpr uint32_t kRowBytes = kItemPos * 4;  // one plane row of a full item
    constexpr uint32_t kStageBytes = PL * kRowBytes;
    extern __shared__ __align__(128) unsigned char bulk_smem[];
    const uint32_t ring = smem_u32(bulk_smem);
    const uint32_t full0 = ring + static_cast<uint32_t>(cfg.stages) * kStageBytes;
    const uint32_t empty0 = full0 + 8u * cfg.stages;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < cfg.stages; ++s) {
            mbar_init(full0 + 8u * s, 1);
            mbar_init(empty0 + 8u * s, CW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int nchunks = (cfg.nq + PL - 1) / PL;
    int stage = 0;
    uint32_t phase = 0;
    if (warp == CW) {
        // ===== producer warp: lane j issues the bulk copy (TMA engine) of plane j of the chunk =====
        for (int it = blockIdx.x; it < total_items; it += gridDim.x) {
            const BulkItem w = bulk_item<kItemPos>(P, cfg, it);
            const uint32_t row_bytes = static_cast<uint32_t>(w.npos) * 4u;
            for (int c = 0; c < nchunks; ++c) {
                const int q0 = c * PL;
                const int nq = min(PL, cfg.nq - q0);
                mbar_wait(empty0 + 8u * stage, phase ^ 1u);
                if (lane == 0) mbar_expect_tx(full0 + 8u * stage, row_bytes * nq);
                __syncwarp();
                if (lane < nq)
                    bulk_g2s(ring + stage * kStageBytes + lane * kRowBytes, bulk_plane(P, w, q0 + lane), row_bytes,
                             full0 + 8u * stage);
                if (++stage == cfg.stages) { stage = 0; phase ^= 1u; }
            }
        }
    } else {
        // ===== consumer warps: thread t owns positions pos0 + 4t .. 4t+3 =====
        const int t = threadIdx.x;
        const int ncls_chunks_full = P.C / PL;  // leading chunks made of class planes only
        for (int it = blockIdx.x; it < total_items; it += gridDim.x) {
            const BulkItem w = bulk_item<kItemPos>(P, cfg, it);
            const bool active = 4 * t < w.npos;  // H*W % 4 == 0: a thread's four positions are all valid or all not
            float m1[4], m2[4], objv[4];
            int k0[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { m1[i] = -INFINITY; m2[i] = -INFINITY; objv[i] = 0.0f; k0[i] = 0; }
            for (int c = 0; c < nchunks; ++c) {
                mbar_wait(full0 + 8u * stage, phase);
                const uint32_t src = ring + stage * kStageBytes + 16u * t;
                if (active) {
                    if (c < ncls_chunks_full) {
#pragma unroll
                        for (int j = 0; j < PL; ++j) {
                            const float4 v = lds128(src + j * kRowBytes);
                            const int k = c * PL + j;
                            top2_update(v.x, k, m1[0], m2[0], k0[0]);
                            top2_update(v.y, k, m1[1], m2[1], k0[1]);
                            top2_update(v.z, k, m1[2], m2[2], k0[2]);
                            top2_update(v.w, k, m1[3], m2[3], k0[3]);
                        }
                    } else {
                        const int q0 = c * PL;
                        const int nq = min(PL, cfg.nq - q0);
                        for (int j = 0; j < nq; ++j) {
                            const float4 v = lds128(src + j * kRowBytes);
                            const int k = q0 + j;
                            if (k == P.C) {
                                objv[0] = v.x; objv[1] = v.y; objv[2] = v.z; objv[3] = v.w;
                            } else {
                                top2_update(v.x, k, m1[0], m2[0], k0[0]);
                                top2_update(v.y, k, m1[1], m2[1], k0[1]);
                                top2_update(v.z, k, m1[2], m2[2], k0[2]);
                                top2_update(v.w, k, m1[3], m2[3], k0[3]);
                            }
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(empty0 + 8u * stage);
                if (++stage == cfg.stages) { stage = 0; phase ^= 1u; }
            }
            const LevelDesc &lv = P.lv[w.l];
            const size_t hw = static_cast<size_t>(lv.hw);
            const float *cbase = lv.p0 + (static_cast<size_t>(w.img * P.A + w.a) * P.cls_nch + P.cls_ch) * hw + w.pos0 + 4 * t;
            uint64_t out[4];
            unsigned okm = 0u;
            int npre = 0;
            uint32_t smax_bits = 0u, smin_inv = 0u;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                out[i] = 0ull;
                if (active) {
                    const float *cj = cbase + i;
                    float score;
                    int cid;
                    bool pre;
                    const bool ok = decide_candidate<false>(
                        P, m1[i], m2[i], k0[i], objv[i], [&](int kk) { return __ldg(cj + static_cast<size_t>(kk) * hw); }, score, cid, pre);
                    npre += pre ? 1 : 0;
                    if (ok) {
                        out[i] = pack_key(score, static_cast<uint32_t>(P.cand_base + lv.cand_off + w.a * lv.hw + w.pos0 + 4 * t + i), static_cast<uint32_t>(cid));
                        okm |= 1u << i;
                        const uint32_t sb = __float_as_uint(score);
                        smax_bits = max(smax_bits, sb);
                        smin_inv = max(smin_inv, ~sb);
                    }
                }
            }
            emit_keys<4>(out, okm, npre, smax_bits, smin_inv, keys + static_cast<int64_t>(w.img) * key_cap, key_cap,
                         counts + w.img * 4, P.pre_kind == PRE_ANY_GT);
        }
    }
}

template <int CW, int PL>
static cudaError_t launch_bulk(const Plan &P, int num_sms, uint64_t *d_keys, int64_t key_cap, int32_t *d_counts, cudaStream_t stream)
{
    constexpr int kItemPos = 128 * CW;
    constexpr int kBulkThreads = (CW + 1) * 32;
    static_assert(PL <= 32, "one producer lane per plane of a stage");
    BulkCfg cfg;
    cfg.nq = P.C + (P.use_obj ? 1 : 0);
    const size_t stage_bytes = static_cast<size_t>(PL) * kItemPos * sizeof(float);
    int stages = static_cast<int>((200 * 1024) / stage_bytes);
    stages = stages > 32 ? 32 : (stages < 2 ? 2 : stages);
    cfg.stages = stages;
    int items = 0;
    for (int l = 0; l < P.L; ++l) {
        cfg.item_off[l] = items;
        cfg.blocks_per_anchor[l] = (P.lv[l].hw + kItemPos - 1) / kItemPos;
        items += P.A * cfg.blocks_per_anchor[l];
    }
    for (int l = P.L; l < YSB_MAX_LEVELS; ++l) { cfg.item_off[l] = items; cfg.blocks_per_anchor[l] = 1; }
    cfg.items_per_img = items;
    const int total = items * P.batch;
    const size_t smem = stage_bytes * stages + 2 * sizeof(uint64_t) * stages;
    {
        cudaError_t e = cudaFuncSetAttribute(k_filter_planes_bulk<CW, PL>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return e;
    }
    const int grid = total < num_sms ? total : num_sms;
    k_filter_planes_bulk<CW, PL><<<grid, kBulkThreads, smem, stream>>>(P, cfg, total, d_keys, key_cap, d_counts);
    return cudaGetLastError();
}

// -------------------------------------------------------------------------------------------------------
// planes layout, 2-D tensor-map TMA version (cp.async.bulk.tensor.2d, SASS UTMALDG).
//
// Each level's head tensor is described to the TMA unit as a 2-D array [rows = batch*A*channels][hw positions]; ONE
// request moves a box of PL consecutive channel rows x up to 256 positions into shared memory, so a ring stage
// (PL planes x 128*CW positions) costs CW/2 requests instead of PL*... per-row bulk copies (the 1-D ring above was
// bound by the producer's request rate).  Out-of-range positions of the last block of a plane are zero-filled by the
// TMA unit without DRAM traffic.  Streamed rows of an (image, anchor): [objectness,] class 0 .. C-1 -- contiguous
// channels for YOLOv5 / YOLOX (objectness is the channel before class 0) and for objectness-free heads (YOLOv8).
// -------------------------------------------------------------------------------------------------------
struct TmaMaps {
    CUtensorMap m[YSB_MAX_LEVELS];
};
struct TmaCfg {
    int stages;
    int nq;          // streamed rows per (image, anchor)
    int obj_first;   // row 0 is the objectness plane
    int row0;        // channel of the first streamed row inside an (image, anchor) block
    int items_per_img;
    int item_off[YSB_MAX_LEVELS];
    int blocks_per_anchor[YSB_MAX_LEVELS];
    int box0[YSB_MAX_LEVELS];  // positions per request on this level: min(256, hw)
};

__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const CUtensorMap *map, int c0, int c1, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst_smem),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar)
                 : "memory");
}

template <int CW, int PL, int BPS>
__global__ void __launch_bounds__((CW + 1) * 32, BPS)
k_filter_planes_tma(const __grid_constant__ Plan P, const __grid_constant__ TmaCfg cfg, const __grid_constant__ TmaMaps maps,
                    int total_items, uint64_t *__restrict__ keys, int64_t key_cap, int32_t *__restrict__ counts)
{
    constexpr int kItemPos = 128 * CW;
    constexpr uint32_t kStageBytes = PL * kItemPos * 4;
    extern __shared__ __align__(128) unsigned char bulk_smem[];
    const uint32_t ring = (smem_u32(bulk_smem) + 127u) & ~127u;
    const uint32_t full0 = ring + static_cast<uint32_t>(cfg.stages) * kStageBytes;
    const uint32_t empty0 = full0 + 8u * cfg.stages;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < cfg.stages; ++s) {
            mbar_init(full0 + 8u * s, 1);
            mbar_init(empty0 + 8u * s, CW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int nchunks = (cfg.nq + PL - 1) / PL;
    int stage = 0;
    uint32_t phase = 0;
    // items are numbered image-fastest: concurrently running CTAs append to different per-image counters
    if (warp == CW) {
        // ===== producer warp: lane r issues request r (256 positions x PL rows) of the stage =====
        for (int it = blockIdx.x; it < total_items; it += gridDim.x) {
            const int img = it % P.batch;
            BulkItem w = bulk_item<kItemPos>(P, cfg, it / P.batch);
            const int box0 = cfg.box0[w.l];
            const int nreq = (w.npos + box0 - 1) / box0;
            const int row_base = (img * P.A + w.a) * P.cls_nch + cfg.row0;
            const uint32_t req_bytes = static_cast<uint32_t>(PL * box0) * 4u;
            for (int c = 0; c < nchunks; ++c) {
                mbar_wait(empty0 + 8u * stage, phase ^ 1u);
                if (lane == 0) mbar_expect_tx(full0 + 8u * stage, req_bytes * nreq);
                __syncwarp();
                if (lane < nreq)
                    tma_load_2d(ring + stage * kStageBytes + lane * req_bytes, &maps.m[w.l], w.pos0 + lane * box0,
                                row_base + c * PL, full0 + 8u * stage);
                if (++stage == cfg.stages) { stage = 0; phase ^= 1u; }
            }
        }
    } else {
        // ===== consumer warps: thread t owns positions pos0 + 4t .. 4t+3 =====
        const int t = threadIdx.x;
        const int nobj = cfg.obj_first;
        for (int it = blockIdx.x; it < total_items; it += gridDim.x) {
            const int img = it % P.batch;
            BulkItem w = bulk_item<kItemPos>(P, cfg, it / P.batch);
            w.img = img;
            const bool active = 4 * t < w.npos;  // H*W % 4 == 0: a thread's four positions are all valid or all not
            const int box0 = cfg.box0[w.l];
            const int req = (4 * t) / box0;
            const uint32_t row_pitch = static_cast<uint32_t>(box0) * 4u;
            const uint32_t my_off = static_cast<uint32_t>(req) * (PL * row_pitch) + static_cast<uint32_t>(4 * t - req * box0) * 4u;
            float m1[4], m2[4], objv[4];
            int k0[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { m1[i] = -INFINITY; m2[i] = -INFINITY; objv[i] = 0.0f; k0[i] = 0; }
            for (int c = 0; c < nchunks; ++c) {
                mbar_wait(full0 + 8u * stage, phase);
                const uint32_t src = ring + stage * kStageBytes + my_off;
                if (active) {
                    const int q0 = c * PL;
                    if (q0 >= nobj && q0 + PL <= cfg.nq) {
#pragma unroll
                        for (int j = 0; j < PL; ++j) {
                            const float4 v = lds128(src + j * row_pitch);
                            const int k = q0 - nobj + j;
                            top2_update(v.x, k, m1[0], m2[0], k0[0]);
                            top2_update(v.y, k, m1[1], m2[1], k0[1]);
                            top2_update(v.z, k, m1[2], m2[2], k0[2]);
                            top2_update(v.w, k, m1[3], m2[3], k0[3]);
                        }
                    } else {
                        const int nq = min(PL, cfg.nq - q0);
#pragma unroll
                        for (int j = 0; j < PL; ++j) {
                            if (j < nq) {
                                const float4 v = lds128(src + j * row_pitch);
                                const int q = q0 + j;
                                if (q < nobj) {
                                    objv[0] = v.x; objv[1] = v.y; objv[2] = v.z; objv[3] = v.w;
                                } else {
                                    const int k = q - nobj;
                                    top2_update(v.x, k, m1[0], m2[0], k0[0]);
                                    top2_update(v.y, k, m1[1], m2[1], k0[1]);
                                    top2_update(v.z, k, m1[2], m2[2], k0[2]);
                                    top2_update(v.w, k, m1[3], m2[3], k0[3]);
                                }
                            }
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(empty0 + 8u * stage);
                if (++stage == cfg.stages) { stage = 0; phase ^= 1u; }
            }
            const LevelDesc &lv = P.lv[w.l];
            const size_t hw = static_cast<size_t>(lv.hw);
            const float *cbase = lv.p0 + (static_cast<size_t>(w.img * P.A + w.a) * P.cls_nch + P.cls_ch) * hw + w.pos0 + 4 * t;
            uint64_t out[4];
            unsigned okm = 0u;
            int npre = 0;
            uint32_t smax_bits = 0u, smin_inv = 0u;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                out[i] = 0ull;
                if (active) {
                    const float *cj = cbase + i;
                    float score;
                    int cid;
                    bool pre;
                    const bool ok = decide_candidate<false>(
                        P, m1[i], m2[i], k0[i], objv[i], [&](int kk) { return __ldg(cj + static_cast<size_t>(kk) * hw); }, score, cid, pre);
                    npre += pre ? 1 : 0;
                    if (ok) {
                        out[i] = pack_key(score, static_cast<uint32_t>(P.cand_base + lv.cand_off + w.a * lv.hw + w.pos0 + 4 * t + i), static_cast<uint32_t>(cid));
                        okm |= 1u << i;
                        const uint32_t sb = __float_as_uint(score);
                        smax_bits = max(smax_bits, sb);
                        smin_inv = max(smin_inv, ~sb);
                    }
                }
            }
            emit_keys<4>(out, okm, npre, smax_bits, smin_inv, keys + static_cast<int64_t>(w.img) * key_cap, key_cap,
                         counts + w.img * 4, P.pre_kind == PRE_ANY_GT);
        }
    }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda at link time)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// the TMA variant covers heads whose streamed channels are contiguous: objectness right before class 0, or none
static bool tma_variant_applies(const Plan &P, int vec)
{
    if (P.layout != LAYOUT_PLANES || vec != 4 || P.multi_label) return false;
    if (P.use_obj && !(P.obj_src == 0 && P.obj_nch == P.cls_nch && P.obj_ch == P.cls_ch - 1)) return false;
    return encode_tiled_fn() != nullptr;
}

template <int CW, int PL, int BPS>
static cudaError_t launch_tma(const Plan &P, int num_sms, uint64_t *d_keys, int64_t key_cap, int32_t *d_counts, cudaStream_t stream)
{
    constexpr int kItemPos = 128 * CW;
    constexpr int kThreads = (CW + 1) * 32;
    static_assert(CW % 2 == 0 && CW / 2 <= 32, "one producer lane per 256-position request");
    TmaCfg cfg;
    TmaMaps maps;
    memset(&maps, 0, sizeof(maps));
    cfg.obj_first = P.use_obj ? 1 : 0;
    cfg.nq = P.C + cfg.obj_first;
    cfg.row0 = P.cls_ch - cfg.obj_first;
    const size_t stage_bytes = static_cast<size_t>(PL) * kItemPos * sizeof(float);
    int stages = static_cast<int>(((216 / BPS) * 1024) / stage_bytes);
    stages = stages > 16 ? 16 : (stages < 2 ? 2 : stages);
    cfg.stages = stages;
    int items = 0;
    EncodeTiledFn enc = encode_tiled_fn();
    for (int l = 0; l < P.L; ++l) {
        const LevelDesc &lv = P.lv[l];
        cfg.item_off[l] = items;
        cfg.blocks_per_anchor[l] = (lv.hw + kItemPos - 1) / kItemPos;
        cfg.box0[l] = lv.hw < 256 ? lv.hw : 256;
        items += P.A * cfg.blocks_per_anchor[l];
        const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(lv.hw), static_cast<cuuint64_t>(P.batch) * P.A * P.cls_nch};
        const cuuint64_t gstride[1] = {static_cast<cuuint64_t>(lv.hw) * sizeof(float)};
        const cuuint32_t box[2] = {static_cast<cuuint32_t>(cfg.box0[l]), static_cast<cuuint32_t>(PL)};
        const cuuint32_t estr[2] = {1, 1};
        const CUresult r = enc(&maps.m[l], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(lv.p0), gdim, gstride, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;
    }
    for (int l = P.L; l < YSB_MAX_LEVELS; ++l) { cfg.item_off[l] = items; cfg.blocks_per_anchor[l] = 1; cfg.box0[l] = 256; }
    cfg.items_per_img = items;
    const int total = items * P.batch;
    const size_t smem = stage_bytes * stages + 2 * sizeof(uint64_t) * stages + 128;
    cudaError_t e = cudaFuncSetAttribute(k_filter_planes_tma<CW, PL, BPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    const int grid = total < num_sms * BPS ? total : num_sms * BPS;
    k_filter_planes_tma<CW, PL, BPS><<<grid, kThreads, smem, stream>>>(P, cfg, maps, total, d_keys, key_cap, d_counts);
    return cudaGetLastError();
}

// -------------------------------------------------------------------------------------------------------
// planes layout, cp.async version.
//
// Persistent CTAs; every thread owns 4 consecutive positions of a work item (128*THREADS/32... = 4*THREADS positions
// of one (image, anchor)) and prefetches ITS OWN 16 bytes of every plane into a private shared-memory ring with
// cp.async (LDGSTS, no register staging).  Because a thread only ever reads bytes it copied itself, the pipeline
// needs no barrier at all: cp.async.wait_group orders the thread's own copies.  NG groups of G planes are kept in
// flight per thread, i.e. (NG-1)*G*16*THREADS bytes per CTA (~180 KB per SM) independent of registers/occupancy.
// -------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void *src)
{
    asm volatile("cp.async.cg.shared.global.L2::128B [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// One "unit" = 4 consecutive positions of one (image, level, anchor) -- the same decomposition as k_filter_planes<4>.
// Threads walk the global unit list with a grid stride, so every SM gets the same mix of work whatever the level
// shapes are; a thread's copy pipeline runs seamlessly from one unit into the next.
struct UnitRef {
    const float *cls;   // class plane 0 at this unit's positions
    const float *obj;   // objectness plane at this unit's positions (nullptr when unused)
    uint32_t hw;        // plane stride in floats
    int cand0;          // candidate index of the first position
    int img;
};

__device__ __forceinline__ UnitRef unit_ref(const Plan &P, int64_t U)
{
    UnitRef r;
    r.img = static_cast<int>(U / P.units_per_img);
    const int u = static_cast<int>(U - static_cast<int64_t>(r.img) * P.units_per_img);
    int l = 0;
#pragma unroll
    for (int i = 1; i < YSB_MAX_LEVELS; ++i)
        if (i < P.L && u >= P.lv[i].unit_off) l = i;
    const LevelDesc &lv = P.lv[l];
    const int upa = lv.hw >> 2;
    const int ru = u - lv.unit_off;
    const int a = ru / upa;
    const int pos = (ru - a * upa) << 2;
    r.hw = static_cast<uint32_t>(lv.hw);
    r.cand0 = lv.cand_off + a * lv.hw + pos;
    r.cls = lv.p0 + (static_cast<size_t>(r.img * P.A + a) * P.cls_nch + P.cls_ch) * lv.hw + pos;
    r.obj = P.use_obj ? (P.obj_src == 2 ? lv.p2 : lv.p0) + (static_cast<size_t>(r.img * P.A + a) * P.obj_nch + P.obj_ch) * lv.hw + pos
                      : nullptr;
    return r;
}

template <int THREADS, int G, int NG>
__global__ void __launch_bounds__(THREADS, (THREADS <= 256 ? 4 : (THREADS <= 512 ? 2 : 1)))
k_filter_planes_async(const __grid_constant__ Plan P, int64_t total_units, uint64_t *__restrict__ keys, int64_t key_cap,
                      int32_t *__restrict__ counts)
{
    constexpr uint32_t kPlaneBytes = 16u * THREADS;  // one plane slot of the ring (all threads of the CTA)
    constexpr uint32_t kGroupBytes = G * kPlaneBytes;
    extern __shared__ __align__(128) unsigned char ring_smem[];
    const int t = threadIdx.x;
    const uint32_t my = smem_u32(ring_smem) + 16u * t;
    const int C = P.C;
    const int nobj = P.use_obj ? 1 : 0;
    const int nq_all = C + nobj;  // streamed planes per unit: [objectness,] class 0 .. C-1
    const int64_t stride = static_cast<int64_t>(gridDim.x) * THREADS;
    const int64_t U0 = static_cast<int64_t>(blockIdx.x) * THREADS + t;

    // ---- issue cursor ------------------------------------------------------------------------------------------
    int64_t i_U = U0;
    int i_q = 0;
    const float *i_src = nullptr;  // address of the next class plane to copy
    const float *i_obj = nullptr;
    uint32_t i_hw = 0;
    auto load_issue_unit = [&]() {
        if (i_U < total_units) {
            const UnitRef r = unit_ref(P, i_U);
            i_src = r.cls;
            i_obj = r.obj;
            i_hw = r.hw;
        }
        i_q = 0;
    };
    load_issue_unit();
    uint32_t i_slot = 0;
    auto issue_one = [&]() {
        if (i_U < total_units) {
            const uint32_t dst = my + i_slot * kGroupBytes;
            if (i_q >= nobj && i_q + G <= nq_all) {  // G class planes
                const float *sp = i_src;
#pragma unroll
                for (int j = 0; j < G; ++j) {
                    cp_async16(dst + j * kPlaneBytes, sp);
                    sp += i_hw;
                }
                i_src = sp;
            } else {
                const int nq = min(G, nq_all - i_q);
                for (int j = 0; j < nq; ++j) {
                    if (i_q + j < nobj) {
                        cp_async16(dst + j * kPlaneBytes, i_obj);
                    } else {
                        cp_async16(dst + j * kPlaneBytes, i_src);
                        i_src += i_hw;
                    }
                }
            }
            i_q += G;
            if (i_q >= nq_all) {
                i_U += stride;
                load_issue_unit();
            }
        }
        cp_async_commit();  // always commit (possibly empty) so that group counting stays uniform
        if (++i_slot == NG) i_slot = 0;
    };

#pragma unroll
    for (int s = 0; s < NG - 1; ++s) issue_one();

    uint32_t c_slot = 0;
    const int lane = t & 31;
    for (int64_t U = U0; U - lane < total_units; U += stride) {  // warp-uniform trip count (emit_keys shuffles)
        const bool valid = U < total_units;
        float m1[4], m2[4], objv[4];
        int k0[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { m1[i] = -INFINITY; m2[i] = -INFINITY; objv[i] = 0.0f; k0[i] = 0; }
        for (int q0 = 0; q0 < nq_all; q0 += G) {
            issue_one();
            cp_async_wait<NG - 1>();
            const uint32_t src = my + c_slot * kGroupBytes;
            if (q0 >= nobj && q0 + G <= nq_all) {
#pragma unroll
                for (int j = 0; j < G; ++j) {
                    const float4 v = lds128(src + j * kPlaneBytes);
                    const int k = q0 - nobj + j;
                    top2_update(v.x, k, m1[0], m2[0], k0[0]);
                    top2_update(v.y, k, m1[1], m2[1], k0[1]);
                    top2_update(v.z, k, m1[2], m2[2], k0[2]);
                    top2_update(v.w, k, m1[3], m2[3], k0[3]);
                }
            } else {
                const int nq = min(G, nq_all - q0);
                for (int j = 0; j < nq; ++j) {
                    const float4 v = lds128(src + j * kPlaneBytes);
                    if (q0 + j < nobj) {
                        objv[0] = v.x; objv[1] = v.y; objv[2] = v.z; objv[3] = v.w;
                    } else {
                        const int k = q0 - nobj + j;
                        top2_update(v.x, k, m1[0], m2[0], k0[0]);
                        top2_update(v.y, k, m1[1], m2[1], k0[1]);
                        top2_update(v.z, k, m1[2], m2[2], k0[2]);
                        top2_update(v.w, k, m1[3], m2[3], k0[3]);
                    }
                }
            }
            if (++c_slot == NG) c_slot = 0;
        }
        uint64_t out[4] = {0ull, 0ull, 0ull, 0ull};
        unsigned okm = 0u;
        int npre = 0, my_img = -1;
        uint32_t smax_bits = 0u, smin_inv = 0u;
        if (valid) {
            const UnitRef r = unit_ref(P, U);
            const size_t hw = r.hw;
            my_img = r.img;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float *cj = r.cls + i;
                float score;
                int cid;
                bool pre;
                const bool ok = decide_candidate<false>(
                    P, m1[i], m2[i], k0[i], objv[i], [&](int kk) { return __ldg(cj + static_cast<size_t>(kk) * hw); }, score, cid, pre);
                npre += pre ? 1 : 0;
                if (ok) {
                    out[i] = pack_key(score, static_cast<uint32_t>(P.cand_base + r.cand0 + i), static_cast<uint32_t>(cid));
                    okm |= 1u << i;
                    const uint32_t sb = __float_as_uint(score);
                    smax_bits = max(smax_bits, sb);
                    smin_inv = max(smin_inv, ~sb);
                }
            }
        }
        // a warp's 32 units may straddle an image boundary: append per image (lane 0 is always valid)
        const int img_lo = __shfl_sync(0xffffffffu, my_img, 0);
        const int img_hi = __reduce_max_sync(0xffffffffu, my_img);
        for (int im = img_lo; im <= img_hi; ++im) {
            const bool mine = my_img == im;
            emit_keys<4>(out, mine ? okm : 0u, mine ? npre : 0, mine ? smax_bits : 0u, mine ? smin_inv : 0u,
                         keys + static_cast<int64_t>(im) * key_cap, key_cap, counts + im * 4, P.pre_kind == PRE_ANY_GT);
        }
    }
    cp_async_wait<0>();
}

template <int THREADS, int G, int NG>
static cudaError_t launch_async(const Plan &P, int num_sms, uint64_t *d_keys, int64_t key_cap, int32_t *d_counts, cudaStream_t stream)
{
    const int64_t total_units = static_cast<int64_t>(P.units_per_img) * P.batch;
    const size_t smem = static_cast<size_t>(NG) * G * 16 * THREADS;
    {
        cudaError_t e = cudaFuncSetAttribute(k_filter_planes_async<THREADS, G, NG>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return e;
    }
    int ctas_per_sm = static_cast<int>((220 * 1024) / (smem + 1024));
    if (ctas_per_sm * THREADS > 2048) ctas_per_sm = 2048 / THREADS;
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    int64_t grid = static_cast<int64_t>(num_sms) * ctas_per_sm;
    const int64_t need = (total_units + THREADS - 1) / THREADS;
    if (grid > need) grid = need;
    k_filter_planes_async<THREADS, G, NG><<<static_cast<unsigned>(grid), THREADS, smem, stream>>>(P, total_units, d_keys, key_cap, d_counts);
    return cudaGetLastError();
}

// -------------------------------------------------------------------------------------------------------
// rows layout (channels-last heads: YOLOv7, RetinaNet cls; and the decoded (b, N, C') tensor of any family).
// A CTA stages a tile of 128 whole rows in shared memory with cp.async (rows are 340 B / 320 B and only 4-byte
// aligned individually, but the tile is contiguous and 16-byte aligned), then one thread scans one row; rows with an
// even float stride are walked with a per-thread rotation so that the 32 rows a warp scans sit in 32 different banks.
// -------------------------------------------------------------------------------------------------------
constexpr int kRowsTile = 128;

__global__ void __launch_bounds__(kRowsTile) k_filter_rows(const __grid_constant__ Plan P, uint64_t *__restrict__ keys,
                                                           int64_t key_cap, int32_t *__restrict__ counts)
{
    extern __shared__ float tile[];
    const int img = blockIdx.x;  // image fastest: see k_filter_planes
    const int t = blockIdx.y;
    int l = 0;
#pragma unroll
    for (int i = 1; i < YSB_MAX_LEVELS; ++i)
        if (i < P.L && t >= P.lv[i].unit_off) l = i;
    const LevelDesc &lv = P.lv[l];
    const int rows_l = (l + 1 < P.L ? P.lv[l + 1].cand_off : P.N) - lv.cand_off;
    const int r0 = (t - lv.unit_off) * kRowsTile;
    const int nrows = min(kRowsTile, rows_l - r0);
    const int rw = P.row_w_in;
    const float *src = lv.p0 + (static_cast<size_t>(img) * lv.img_rows + r0) * rw;
    const int nfl = nrows * rw;
    // 16-byte aligned tiles (every reference layout): cp.async straight into a DENSE shared tile -- no register
    // staging, no per-element address math.  Rows with an even float stride are scanned with a per-thread rotation
    // (below) instead of padding.  Odd shapes: scalar loads into a tile padded to an odd stride.
    const bool dense = ((reinterpret_cast<uintptr_t>(src) & 15u) == 0) && ((nfl & 3) == 0);
    const int sstride = dense ? rw : (rw | 1);
    if (dense) {
        const uint32_t dst = smem_u32(tile);
        const int nv = nfl >> 2;
        for (int i = threadIdx.x; i < nv; i += kRowsTile) cp_async16(dst + 16u * i, src + 4 * i);
        cp_async_commit();
        cp_async_wait<0>();
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
        // thread t reads word sstride*t + k: conflict-free when sstride is odd; otherwise start the walk at column t so
        // that the 32 lanes hit (sstride+1)*t + k.  The visiting order is irrelevant: a unique maximum has a unique
        // index, and equal maxima force m2 == m1, i.e. the literal path, which walks in index order.
        if (sstride & 1) {
            // odd stride (YOLOv7's 85-float rows): already conflict-free, a straight run without the wrap test
#pragma unroll 8
            for (int k = 0; k < P.C; ++k) top2_update(cls[k], k, m1, m2, k0);
        } else {
            // rotated walk: lane l starts at column l, i.e. reads word sstride*tid + l + t = (sstride + 1)*l + t (mod 32)
            // -- an odd multiplier, conflict-free.  Every lane runs the same C iterations (two straight runs per lane
            // would diverge); the first C - 32 of them cannot wrap, so only the last 32 carry the wrap test.
            const int lane = static_cast<int>(threadIdx.x) & 31;
            if (P.C >= 32) {
                const int straight = P.C - 32;
#pragma unroll 8
                for (int t = 0; t < straight; ++t) top2_update(cls[t + lane], t + lane, m1, m2, k0);
                int k = straight + lane;
#pragma unroll 8
                for (int t = 0; t < 32; ++t) {
                    if (k >= P.C) k -= P.C;
                    top2_update(cls[k], k, m1, m2, k0);
                    ++k;
                }
            } else {
                int k = lane % P.C;
                for (int t = 0; t < P.C; ++t) {
                    top2_update(cls[k], k, m1, m2, k0);
                    if (++k == P.C) k = 0;
                }
            }
        }
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
            out[0] = pack_key(score, static_cast<uint32_t>(P.cand_base + cand), static_cast<uint32_t>(c));
            okm = 1u;
            smax_bits = __float_as_uint(score);
            smin_inv = ~smax_bits;
        }
    }
    emit_keys<1>(out, okm, npre, smax_bits, smin_inv, keys + static_cast<int64_t>(img) * key_cap, key_cap,
                 counts + img * 4, P.pre_kind == PRE_ANY_GT);
}

// -------------------------------------------------------------------------------------------------------
// mutil_label: true -- one record per (candidate, class) whose score reaches the class threshold
// (trainer/eval_yolov5.py:276-279, eval_yolov7.py:230-233, eval_yolox.py:217-220, eval_yolov8.py:184-187,
// eval_fcos.py:246-250).  False in every shipped yaml, so this kernel favours simplicity: one thread per candidate,
// every class evaluated literally (sigmoid, multiply) twice -- once to count, once to write -- any input layout.
// -------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_filter_multilabel(const __grid_constant__ Plan P, uint64_t *__restrict__ keys,
                                                           int64_t key_cap, int32_t *__restrict__ counts)
{
    const int img = blockIdx.x;
    const int cand = blockIdx.y * blockDim.x + threadIdx.x;
    const bool active = cand < P.N;
    const bool decoded = P.input_kind == YSB_INPUT_DECODED_ROWS;
    const float *cls = nullptr;
    size_t cstride = 1;
    float mult = 1.0f;
    if (active) {
        const int l = find_level(P, cand);
        const LevelDesc &lv = P.lv[l];
        const int r = cand - lv.cand_off;
        float objv = 0.0f;
        if (P.layout == LAYOUT_PLANES) {
            const int a = r / lv.hw, pos = r - a * lv.hw;
            cls = lv.p0 + (static_cast<size_t>(img * P.A + a) * P.cls_nch + P.cls_ch) * lv.hw + pos;
            cstride = static_cast<size_t>(lv.hw);
            if (P.use_obj)
                objv = __ldg((P.obj_src == 2 ? lv.p2 : lv.p0) + (static_cast<size_t>(img * P.A + a) * P.obj_nch + P.obj_ch) * lv.hw + pos);
        } else {
            const float *row = lv.p0 + (static_cast<size_t>(img) * lv.img_rows + r) * P.row_w_in;
            cls = row + P.cls_col_in;
            if (P.use_obj)
                objv = P.obj_src == 1 ? __ldg(P.lv[0].p1 + (static_cast<size_t>(img) * P.N + cand) * P.reg_row_w + 4)
                                      : __ldg(row + P.obj_col_in);
        }
        if (P.use_obj) mult = decoded ? objv : sigmoid_ref(objv);
    }
    auto score_of = [&](int k, float &sk) {
        const float c = __ldg(cls + static_cast<size_t>(k) * cstride);
        sk = decoded ? c : sigmoid_ref(c);
        return P.use_obj ? __fmul_rn(sk, mult) : sk;
    };
    auto passes = [&](float p) { return P.multi_strict ? (p > P.cls_thr) : (p >= P.cls_thr); };
    int n = 0, npre = 0;
    uint32_t smax_bits = 0u, smin_inv = 0u;
    if (active) {
        float smax = -INFINITY, pmax = -INFINITY;
        for (int k = 0; k < P.C; ++k) {
            float sk;
            const float p = score_of(k, sk);
            smax = fmaxf(smax, sk);
            pmax = fmaxf(pmax, p);
            if (sk > P.pre_thr) ++npre;
            if (passes(p)) {
                ++n;
                const uint32_t sb = __float_as_uint(p);
                smax_bits = max(smax_bits, sb);
                smin_inv = max(smin_inv, ~sb);
            }
        }
        bool pre = true;
        switch (P.pre_kind) {
        case PRE_OBJ: pre = mult >= P.conf_thr; break;
        case PRE_OBJ_X_MAX: pre = pmax >= P.conf_thr; break;
        case PRE_MAXCLS: pre = smax >= P.cls_thr; break;
        case PRE_ANY_GT: pre = smax > P.pre_thr; break;
        default: break;
        }
        if (!pre) { n = 0; smax_bits = 0u; smin_inv = 0u; }
        if (P.pre_kind != PRE_ANY_GT) npre = 0;
    }
    // warp-aggregated reservation of n slots per thread
    const unsigned lane = threadIdx.x & 31u;
    int incl = n;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= static_cast<unsigned>(d)) incl += v;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    int32_t *cnt = counts + img * 4;
    const int tp = __reduce_add_sync(0xffffffffu, npre);
    if (lane == 0 && tp) atomicAdd(cnt + 1, tp);  // FCOS: pairs above pre_nms_thresh (eval_fcos.py:246)
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
    if (n > 0) {
        int64_t at = static_cast<int64_t>(base) + (incl - n);
        uint64_t *dst = keys + static_cast<int64_t>(img) * key_cap;
        for (int k = 0; k < P.C; ++k) {
            float sk;
            const float p = score_of(k, sk);
            if (passes(p)) {
                if (at < key_cap) dst[at] = pack_key(p, static_cast<uint32_t>(P.cand_base + cand), static_cast<uint32_t>(k));
                ++at;
            }
        }
    }
}

// -------------------------------------------------------------------------------------------------------
// host launchers
// -------------------------------------------------------------------------------------------------------
// 1 = direct 128-bit loads (default: fastest measured, profiles/README.md), 0 = cp.async private rings,
// 2 = TMA bulk-copy ring.  YSB_FILTER_VARIANT / YSB_BULK_PPT are profiling switches only.
static int g_filter_variant = [] {
    const char *v = getenv("YSB_FILTER_VARIANT");
    return v ? atoi(v) : 1;
}();
static int g_bulk_ppt = [] {
    const char *v = getenv("YSB_BULK_PPT");
    return v ? atoi(v) : 0;
}();

// zero_counts = false: append to the key lists / counters a previous pass left (test-time-augmentation passes)
cudaError_t launch_filter(const Plan &P, int vec, uint64_t *d_keys, int64_t key_cap, int32_t *d_counts,
                          cudaStream_t stream, bool zero_counts)
{
    cudaError_t e = cudaSuccess;
    if (zero_counts) e = cudaMemsetAsync(d_counts, 0, sizeof(int32_t) * 4 * static_cast<size_t>(P.batch), stream);
    if (e != cudaSuccess) return e;
    if (P.batch == 0 || P.N == 0) return cudaSuccess;
    if (P.multi_label) {
        const dim3 grid(P.batch, (P.N + 255) / 256);
        k_filter_multilabel<<<grid, 256, 0, stream>>>(P, d_keys, key_cap, d_counts);
        return cudaGetLastError();
    }
    if (g_filter_variant == 3 && tma_variant_applies(P, vec)) {  // 2-D tensor-map TMA ring
        int num_sms = 0, dev = 0;
        e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        e = cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
        const bool nine = (P.C + (P.use_obj ? 1 : 0)) % 9 == 0;
        switch (g_bulk_ppt) {
        case 4: return nine ? launch_tma<4, 9, 2>(P, num_sms, d_keys, key_cap, d_counts, stream) : launch_tma<4, 8, 2>(P, num_sms, d_keys, key_cap, d_counts, stream);
        case 2: return nine ? launch_tma<2, 9, 4>(P, num_sms, d_keys, key_cap, d_counts, stream) : launch_tma<2, 8, 4>(P, num_sms, d_keys, key_cap, d_counts, stream);
        case 163: return nine ? launch_tma<16, 3, 1>(P, num_sms, d_keys, key_cap, d_counts, stream) : launch_tma<16, 4, 1>(P, num_sms, d_keys, key_cap, d_counts, stream);
        case 83: return nine ? launch_tma<8, 3, 2>(P, num_sms, d_keys, key_cap, d_counts, stream) : launch_tma<8, 4, 2>(P, num_sms, d_keys, key_cap, d_counts, stream);
        case 43: return nine ? launch_tma<4, 3, 4>(P, num_sms, d_keys, key_cap, d_counts, stream) : launch_tma<4, 4, 4>(P, num_sms, d_keys, key_cap, d_counts, stream);
        case 16: return nine ? launch_tma<16, 9, 1>(P, num_sms, d_keys, key_cap, d_counts, stream) : launch_tma<16, 8, 1>(P, num_sms, d_keys, key_cap, d_counts, stream);
        default: return nine ? launch_tma<8, 9, 1>(P, num_sms, d_keys, key_cap, d_counts, stream) : launch_tma<8, 8, 1>(P, num_sms, d_keys, key_cap, d_counts, stream);
        }
    }
    if (P.layout == LAYOUT_PLANES && vec == 4 && g_filter_variant != 1 && g_filter_variant != 3) {  // ring variants need every level vectorised
        int num_sms = 0, dev = 0;
        e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        e = cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
        const int nq = P.C + (P.use_obj ? 1 : 0);
        const bool nine = nq % 9 == 0 || P.C % 9 == 0;  // e.g. 80 classes + objectness = 9 x 9 planes
        if (g_filter_variant == 2) {  // TMA bulk-copy ring (profiling variant)
            switch (g_bulk_ppt) {
            case 8: e = nine ? launch_bulk<8, 9>(P, num_sms, d_keys, key_cap, d_counts, stream) : launch_bulk<8, 8>(P, num_sms, d_keys, key_cap, d_counts, stream); break;
            default: e = nine ? launch_bulk<4, 9>(P, num_sms, d_keys, key_cap, d_counts, stream) : launch_bulk<4, 8>(P, num_sms, d_keys, key_cap, d_counts, stream); break;
            }
            return e;
        }
        switch (g_bulk_ppt) {  // THREADS, planes per group, groups in the ring
        case 1: e = launch_async<1024, 4, 3>(P, num_sms, d_keys, key_cap, d_counts, stream); break;
        case 2: e = launch_async<1024, 5, 2>(P, num_sms, d_keys, key_cap, d_counts, stream); break;
        case 3: e = launch_async<768, 8, 2>(P, num_sms, d_keys, key_cap, d_counts, stream); break;
        case 4: e = launch_async<512, 4, 3>(P, num_sms, d_keys, key_cap, d_counts, stream); break;
        case 5: e = launch_async<512, 5, 2>(P, num_sms, d_keys, key_cap, d_counts, stream); break;
        case 6: e = launch_async<256, 4, 3>(P, num_sms, d_keys, key_cap, d_counts, stream); break;
        case 7: e = launch_async<256, 8, 3>(P, num_sms, d_keys, key_cap, d_counts, stream); break;
        default: e = launch_async<512, 8, 3>(P, num_sms, d_keys, key_cap, d_counts, stream); break;
        }
        return e;
    } else if (P.layout == LAYOUT_PLANES) {
        static const bool img_fast = getenv("YSB_IMG_FAST") ? atoi(getenv("YSB_IMG_FAST")) != 0 : true;
        const dim3 grid = img_fast ? dim3(P.batch, (P.units_per_img + 255) / 256) : dim3((P.units_per_img + 255) / 256, P.batch);
        const dim3 grid128 = img_fast ? dim3(P.batch, (P.units_per_img + 127) / 128) : dim3((P.units_per_img + 127) / 128, P.batch);
        if (vec == 1) {
            k_filter_planes<1><<<grid, 256, 0, stream>>>(P, d_keys, key_cap, d_counts);
        } else if (vec == 4 && (g_bulk_ppt == 0 || (g_bulk_ppt >= 30 && g_bulk_ppt <= 45))) {
            // every level 128-bit loadable (the common case): specialised kernel; g_bulk_ppt 30..45 = its tuning variants
            const dim3 gv4(P.batch, (P.units_per_img + 127) / 128);
            switch (g_bulk_ppt) {
            case 31: k_filter_planes_v4<16, 128, 4><<<gv4, 128, 0, stream>>>(P, d_keys, key_cap, d_counts); break;
            case 33: k_filter_planes_v4<9, 128, 6><<<gv4, 128, 0, stream>>>(P, d_keys, key_cap, d_counts); break;
            case 37: k_filter_planes_v4<6, 128, 8><<<gv4, 128, 0, stream>>>(P, d_keys, key_cap, d_counts); break;
            case 41: k_filter_planes_v4<9, 64, 12><<<dim3(P.batch, (P.units_per_img + 63) / 64), 64, 0, stream>>>(P, d_keys, key_cap, d_counts); break;
            case 30: k_filter_planes_v4<12, 128, 5><<<gv4, 128, 0, stream>>>(P, d_keys, key_cap, d_counts); break;
            // default: 9 loads of 128 bits in flight per thread (YOLOv5/YOLOX: 80 classes + objectness = 9 batches of 9),
            // 76 registers, 6 CTAs of 128 threads per SM
            default: k_filter_planes_v4<9, 128, 6><<<gv4, 128, 0, stream>>>(P, d_keys, key_cap, d_counts); break;
            }
        } else {
            switch (g_bulk_ppt) {  // profiling variants of the direct-load kernel
            case 3: k_filter_planes<4, 4, 256, 6><<<grid, 256, 0, stream>>>(P, d_keys, key_cap, d_counts); break;
            case 4: k_filter_planes<4, 8, 256, 1, 1><<<grid, 256, 0, stream>>>(P, d_keys, key_cap, d_counts); break;
            case 5: k_filter_planes<4, 16, 256, 2><<<grid, 256, 0, stream>>>(P, d_keys, key_cap, d_counts); break;
            case 6: k_filter_planes<4, 8, 128, 8><<<grid128, 128, 0, stream>>>(P, d_keys, key_cap, d_counts); break;
            case 7: k_filter_planes<4, 10, 256, 3><<<grid, 256, 0, stream>>>(P, d_keys, key_cap, d_counts); break;
            case 8: k_filter_planes<4, 10, 256, 3, 1><<<grid, 256, 0, stream>>>(P, d_keys, key_cap, d_counts); break;
            case 9: k_filter_planes<4, 10, 128, 6, 1><<<grid128, 128, 0, stream>>>(P, d_keys, key_cap, d_counts); break;
            case 10: k_filter_planes<4, 10, 128, 6><<<grid128, 128, 0, stream>>>(P, d_keys, key_cap, d_counts); break;
            case 11: k_filter_planes<4, 20, 128, 4><<<grid128, 128, 0, stream>>>(P, d_keys, key_cap, d_counts); break;
            case 1: k_filter_planes<4><<<grid, 256, 0, stream>>>(P, d_keys, key_cap, d_counts); break;
            case 20: k_filter_planes<4, 16, 128, 4, 1, 1><<<grid128, 128, 0, stream>>>(P, d_keys, key_cap, d_counts); break;  // load-pattern probe
            case 21: k_filter_planes<4, 8, 256, 4, 1><<<grid, 256, 0, stream>>>(P, d_keys, key_cap, d_counts); break;
            case 22: k_filter_planes<4, 8, 128, 8, 1><<<grid128, 128, 0, stream>>>(P, d_keys, key_cap, d_counts); break;
            case 24: k_filter_planes<4, 16, 128, 4, 1><<<grid128, 128, 0, stream>>>(P, d_keys, key_cap, d_counts); break;
            // generic kernel (levels that are not all 128-bit loadable, e.g. FCOS' 5x5 map; YSB_BULK_PPT=50 forces it):
            // 128-thread CTAs, 12 loads of 128 bits in flight per thread (96 registers, 5 CTAs/SM), 256-byte L2 prefetch
            default: k_filter_planes<4, 12, 128, 5, 1><<<grid128, 128, 0, stream>>>(P, d_keys, key_cap, d_counts); break;
            }
        }
    } else {
        const size_t smem = static_cast<size_t>(kRowsTile) * (P.row_w_in | 1) * sizeof(float);
        if (smem > 48 * 1024) {
            e = cudaFuncSetAttribute(k_filter_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
            if (e != cudaSuccess) return e;
        }
        const dim3 grid(P.batch, P.units_per_img);
        k_filter_rows<<<grid, kRowsTile, smem, stream>>>(P, d_keys, key_cap, d_counts);
    }
    return cudaGetLastError();
}

}  // namespace ysb

// filter_variants.cuh -- PROFILING-ONLY variants of the planes-layout filter kernel (compiled only with
// -DYSB_PROFILING_VARIANTS; never part of the product library): a cp.async private-ring kernel, a 1-D TMA bulk-copy
// (cp.async.bulk + mbarrier) warp-specialised ring and a 2-D tensor-map TMA ring.  Same survivor sets as the product
// kernel; all measured slower on this access pattern (profiles/README.md).  Selected at run time by YSB_FILTER_VARIANT /
// YSB_BULK_PPT in such a build.  Included by filter_kernels.cu inside namespace ysb, after the shared helpers.
#pragma once
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
// (mbar_init / mbar_expect_tx / mbar_arrive / mbar_wait / bulk_g2s: filter_kernels.cu, shared with k_filter_rows_ring)
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


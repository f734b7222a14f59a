// nms_kernel.cu -- K2: per-image top-k selection + sort + greedy class-aware NMS + box post-filter.
//
// Replaces, per image, trainer/eval_yolov5.py:293-316 (class offset, utils.nms.numba_nms, max_det truncation,
// postprocess_bbox count filter, row gather) and the same tail in the other evaluators; the NMS itself is
// utils/nms.py:10-27 with the IoU of utils/bbox_tools.py:12-35.
//
// The reference runs the greedy loop over all M survivors (O(K*M), ~14 s per image at M = 25 k) and truncates to
// max_det afterwards.  Greedy NMS is prefix-stable, so this kernel visits candidates in the reference's order
// (score desc, candidate asc) and stops at max_det keeps:
//   1. radix-select (shared-memory histograms over the normalised 64-bit keys) the next <= 4096 best keys,
//   2. bitonic-sort them in shared memory, decode their boxes (4 channels each) from the head tensors,
//   3. walk them in chunks of 64: test each against the kept list (parallel), build the 64x64 in-chunk
//      suppression bitmask (parallel), resolve the chunk with a bitmask sweep (one warp), append the keeps,
//   4. if fewer than max_det boxes are kept and candidates remain, select the next tranche and continue,
//   5. postprocess_bbox count filter / RetinaNet merge / remove_small_boxes, ordered write of the rows.
// One CTA (1024 threads) per image; images of a batch run concurrently on different SMs.
#include "ysb_internal.cuh"

namespace ysb {

constexpr int kThreads = 1024;
constexpr int kTrancheCap = 4096;
constexpr int kMinTranche = 512;
constexpr int kDigitBits = 11;
constexpr int kBins = 1 << kDigitBits;
constexpr int kChunk = 64;
constexpr int kMaxKeep = YSB_MAX_DET_LIMIT;
constexpr int kKeyBatch = 8;  // independent 64-bit key loads in flight per thread in the selection passes

// Phase timestamps of image 0..63 (profiling builds only: -DYSB_K2_TIMING), read back by ysb_debug_k2_timing().
#ifdef YSB_K2_TIMING
__device__ long long g_k2_timing[64][16];
#define K2_STAMP(slot) do { if (threadIdx.x == 0 && blockIdx.x < 64) g_k2_timing[blockIdx.x][slot] = clock64(); } while (0)
#define K2_ACC_BEGIN() k2_t0 = clock64()
#define K2_ACC(slot) do { const long long k2_t1 = clock64(); k2_acc[slot - 8] += k2_t1 - k2_t0; k2_t0 = k2_t1; } while (0)
#define K2_ACC_RESET() long long k2_acc[5] = {0, 0, 0, 0, 0}; long long k2_t0 = 0
#define K2_ACC_FLUSH() do { if (threadIdx.x == 0 && blockIdx.x < 64) for (int q = 0; q < 5; ++q) g_k2_timing[blockIdx.x][8 + q] = k2_acc[q]; } while (0)
#else
#define K2_ACC_BEGIN() do { } while (0)
#define K2_ACC(slot) do { } while (0)
#define K2_ACC_RESET() do { } while (0)
#define K2_ACC_FLUSH() do { } while (0)
#define K2_STAMP(slot) do { } while (0)
#endif

struct NmsSmem {
    uint64_t keys[kTrancheCap];
    float4 raw[kTrancheCap];
    uint32_t hist[kBins];
    float2 kept_x[kMaxKeep];
    float2 kept_y[kMaxKeep];
    float kept_a[kMaxKeep];
    uint64_t kept_key[kMaxKeep];
    float4 kept_raw[kMaxKeep];
    uint8_t kept_flag[kMaxKeep];
    float2 chunk_x[kChunk];
    float2 chunk_y[kChunk];
    float chunk_a[kChunk];
    uint64_t chunk_mask[kChunk];   // row i: later candidates j > i that i suppresses
    unsigned int chunk_pred[kChunk][2];  // row i: earlier candidates j < i that suppress i (lo/hi words)
    float area[kTrancheCap];       // post-filter: areas of the offset boxes
    uint8_t chunk_alive[kChunk];
    uint32_t warp_tmp[kThreads / 32];
    unsigned long long sel_lo;
    unsigned long long keep_mask;
    int n_sel;
    int sel_digit;
    int sel_above;
    int out_count;
};

// Arguments of the array flavour (utils.numba_nms / utils.gpu_nms on one explicit box array).
struct ArrayArgs {
    const float4 *boxes;   // (m) xyxy
    BoxSoA kept;           // global workspace (m boxes): the keep list is unbounded here
    int32_t *keep;         // out: kept indices, visiting order
    int32_t *keep_cnt;     // out
    int iou_kind, cmp;
    float thr32;           // torch compares float32 IoUs against the threshold rounded to float32
};

template <bool ARRAY>
__device__ __forceinline__ bool pair_hit(const BoxSoA &list, int i, const OffBox &b, const IouThr &t, const ArrayArgs &aa)
{
    if (!ARRAY) return iou_reaches_staged<false>(list, i, b, t);
    if (aa.iou_kind == YSB_IOU_NUMBA_F64MIX)
        return aa.cmp == YSB_CMP_GT ? iou_reaches_staged<true>(list, i, b, t) : iou_reaches_staged<false>(list, i, b, t);
    const OffBox a = soa_load(list, i);
    const float v = iou_kind_f32(aa.iou_kind, make_float4(a.x1, a.y1, a.x2, a.y2), make_float4(b.x1, b.y1, b.x2, b.y2));
    return aa.cmp == YSB_CMP_GT ? (v > aa.thr32) : (v >= aa.thr32);
}

__device__ __forceinline__ uint64_t norm_key(uint64_t key, uint32_t smin)
{
    return (static_cast<uint64_t>(static_cast<uint32_t>(key >> 32) - smin) << 32) | (key & 0xffffffffull);
}

// Block-wide radix descent over the normalised keys: returns a lower bound `lo` such that the set
// {lo <= nk <= hi_incl} holds at least kMinTranche keys (or everything that is left) and at most kTrancheCap.
// The smallest such set at the coarsest digit that resolves it is taken: the NMS walk usually ends long before a
// tranche is exhausted, so sorting more than it needs is wasted work.
__device__ uint64_t select_lower_bound(NmsSmem &S, const uint64_t *__restrict__ keys, int M, uint32_t smin,
                                       uint64_t hi_incl, int nbits)
{
    const int tid = threadIdx.x;
    int sh = nbits > kDigitBits ? nbits - kDigitBits : 0;  // shift of the current digit
    int width = nbits - sh;                                // bits in the current digit
    uint64_t prefix = 0;                                   // value of nk >> (sh + width) along the descent path
    bool have_prefix = false;
    int acc = 0;                                           // keys already covered above the path
    for (;;) {
        for (int i = tid; i < kBins; i += kThreads) S.hist[i] = 0;
        if (tid == 0) { S.sel_digit = -1; S.n_sel = 0; S.sel_above = 0; }
        __syncthreads();
        const int top = sh + width;
        const uint32_t dmask = (1u << width) - 1u;
        for (int i0 = 0; i0 < M; i0 += kKeyBatch * kThreads) {
            uint64_t kk[kKeyBatch];
#pragma unroll
            for (int q = 0; q < kKeyBatch; ++q) {
                const int i = i0 + q * kThreads + tid;
                kk[q] = i < M ? __ldg(keys + i) : 0ull;
            }
#pragma unroll
            for (int q = 0; q < kKeyBatch; ++q) {
                if (i0 + q * kThreads + tid >= M) continue;
                const uint64_t nk = norm_key(kk[q], smin);
                if (nk <= hi_incl && (!have_prefix || (nk >> top) == prefix))
                    atomicAdd(&S.hist[static_cast<uint32_t>(nk >> sh) & dmask], 1u);
            }
        }
        __syncthreads();
        // suffix sums S(d) = #keys with digit >= d; thread t owns digits 2t, 2t+1
        const uint32_t h0 = S.hist[2 * tid], h1 = S.hist[2 * tid + 1];
        const uint32_t part = h0 + h1;
        uint32_t incl = part;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_down_sync(0xffffffffu, incl, d);
            if ((tid & 31) + d < 32) incl += v;
        }
        if ((tid & 31) == 0) S.warp_tmp[tid >> 5] = incl;
        __syncthreads();
        uint32_t above = 0;
        for (int w = (tid >> 5) + 1; w < kThreads / 32; ++w) above += S.warp_tmp[w];
        const uint32_t s_hi = incl - part + above;  // S(2t+2)
        const uint32_t s1 = s_hi + h1;              // S(2t+1)
        const uint32_t s0 = s1 + h0;                // S(2t)
        const uint32_t need = static_cast<uint32_t>(kMinTranche > acc ? kMinTranche - acc : 0);
        // d_a = largest digit with S(d_a) >= need
        if (s1 >= need && s_hi < need) { S.sel_digit = 2 * tid + 1; S.n_sel = static_cast<int>(s1); S.sel_above = static_cast<int>(s_hi); }
        else if (s0 >= need && s1 < need) { S.sel_digit = 2 * tid; S.n_sel = static_cast<int>(s0); S.sel_above = static_cast<int>(s1); }
        __syncthreads();
        const int da = S.sel_digit, covered = acc + S.n_sel, abv = acc + S.sel_above;
        __syncthreads();
        const uint64_t base = have_prefix ? (prefix << top) : 0ull;
        if (da < 0) return base;  // fewer than kMinTranche keys under this prefix: take them all
        if (covered <= kTrancheCap || sh == 0) return base | (static_cast<uint64_t>(da) << sh);
        // digit d_a alone overflows the tranche: refine inside it
        prefix = (have_prefix ? (prefix << width) : 0ull) | static_cast<uint64_t>(da);
        have_prefix = true;
        acc = abv;
        const int nw = sh < kDigitBits ? sh : kDigitBits;
        sh -= nw;
        width = nw;
    }
}

// Descending sort of keys[0..n) (n <= 1024*E) by a bitonic network held in registers: element i = tid + 1024*r lives in
// register r of thread tid.  Strides >= 1024 are thread-local, strides < 32 use warp shuffles, only strides 32..512 go
// through shared memory (15 of the 55 stages at n = 1024).  Slots beyond n sort as 0 (smaller than any real key).
template <int E>
__device__ void bitonic_sort_desc(uint64_t *keys, int n)
{
    const int tid = threadIdx.x;
    constexpr int n2 = kThreads * E;
    uint64_t v[E];
#pragma unroll
    for (int r = 0; r < E; ++r) {
        const int i = tid + kThreads * r;
        v[r] = i < n ? keys[i] : 0ull;
    }
    __syncthreads();
    for (int k = 2; k <= n2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= kThreads) {
                // thread-local stage: register pairs (0,1),(2,3) for stride 1024, (0,2),(1,3) for stride 2048
                auto cx = [&](int r0, uint64_t &a, uint64_t &b) {
                    const bool desc = ((tid + kThreads * r0) & k) == 0;
                    const uint64_t hi = a > b ? a : b, lo = a > b ? b : a;
                    a = desc ? hi : lo;
                    b = desc ? lo : hi;
                };
                if (E >= 2 && j == kThreads) {
                    cx(0, v[0], v[1 % E]);
                    if (E == 4) cx(2, v[2 % E], v[3 % E]);
                } else if (E == 4 && j == 2 * kThreads) {
                    cx(0, v[0], v[2 % E]);
                    cx(1, v[1 % E], v[3 % E]);
                }
            } else if (j >= 32) {
#pragma unroll
                for (int r = 0; r < E; ++r) keys[tid + kThreads * r] = v[r];
                __syncthreads();
#pragma unroll
                for (int r = 0; r < E; ++r) {
                    const int i = tid + kThreads * r;
                    const uint64_t o = keys[i ^ j];
                    const bool take_max = ((i & j) == 0) == ((i & k) == 0);
                    v[r] = take_max ? (v[r] > o ? v[r] : o) : (v[r] > o ? o : v[r]);
                }
                __syncthreads();
            } else {
#pragma unroll
                for (int r = 0; r < E; ++r) {
                    const int i = tid + kThreads * r;
                    const uint64_t o = __shfl_xor_sync(0xffffffffu, v[r], j);
                    const bool take_max = ((i & j) == 0) == ((i & k) == 0);
                    v[r] = take_max ? (v[r] > o ? v[r] : o) : (v[r] > o ? o : v[r]);
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < E; ++r) keys[tid + kThreads * r] = v[r];
    __syncthreads();
}

// order-preserving float <-> int maps (signed integer compare == float compare; used for shared-memory atomicMin/Max)
__device__ __forceinline__ int float_ordered(float f)
{
    const int i = __float_as_int(f);
    return i ^ ((i >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float ordered_float(int i) { return __int_as_float(i ^ ((i >> 31) & 0x7fffffff)); }

// Plan a merged candidate index belongs to (test-time-augmentation passes); cand becomes the index inside that pass.
__device__ __forceinline__ const Plan &pass_of(const Plan &P, const NoExtraPasses &, int &) { return P; }
__device__ __forceinline__ const Plan &pass_of(const Plan &P, const ExtraPasses &X, int &cand)
{
    int sel = -1;
#pragma unroll
    for (int i = 0; i < YSB_MAX_PASSES - 1; ++i)
        if (i < X.n && cand >= X.base[i]) sel = i;
    if (sel < 0) return P;
    cand -= X.base[sel];
    return X.p[sel];
}

// EXTRA = NoExtraPasses (one set of heads) or ExtraPasses (TTA: boxes are decoded from the pass a candidate came from;
// thresholds / NMS settings are those of P, identical in every pass).
template <bool ARRAY, typename EXTRA>
__global__ void __launch_bounds__(kThreads, 1)
k_select_nms(const __grid_constant__ Plan P, const __grid_constant__ EXTRA X, const uint64_t *__restrict__ keys_all,
             int64_t key_cap, const int32_t *__restrict__ counts, float *__restrict__ dets,
             int32_t *__restrict__ det_idx, int32_t *__restrict__ det_cnt, const ArrayArgs aa,
             const __grid_constant__ GatherSink G)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    NmsSmem &S = *reinterpret_cast<NmsSmem *>(smem_raw);
    const int tid = threadIdx.x;
    const int img = blockIdx.x;
    const int32_t *cnt = counts + img * 4;
    const int64_t m64 = cnt[0];
    const int M = static_cast<int>(m64 < key_cap ? m64 : key_cap);
    if (M <= 0) {
        if (tid == 0) {
            if (ARRAY) *aa.keep_cnt = 0;
            else if (G.world == 0) det_cnt[img] = -1;  // no survivor: the reference appends None
        }
        if (!ARRAY && G.world > 0 && tid < G.world) {   // detection gather: the count goes to every rank's slot
            G.cnt[tid][img] = -1;
            __threadfence_system();
            atomicAdd_system(G.arrived[tid], 1u);
        }
        return;
    }
    const BoxSoA kept_box = ARRAY ? aa.kept : BoxSoA{S.kept_x, S.kept_y, S.kept_a};
    const BoxSoA chunk_box{S.chunk_x, S.chunk_y, S.chunk_a};
    const uint64_t *keys = keys_all + static_cast<int64_t>(img) * key_cap;
    const uint32_t smax = static_cast<uint32_t>(cnt[2]);
    const uint32_t smin = ~static_cast<uint32_t>(cnt[3]);
    const int nbits = 32 + (smax > smin ? 32 - __clz(smax - smin) : 0);
    const int limit = P.topk_sqrt ? min(M, min(cnt[1], P.pre_nms_topk)) : M;
    const IouThr thr = make_iou_thr(P.iou_thr);
    const int max_det = P.max_det;
    // postprocess_bbox runs only for 1 < M < 3000 (FCOS: <= 300), trainer/eval_yolov5.py:306-307
    const bool window = !ARRAY && P.postprocess_bbox && M > 1 && M < P.window_hi && M <= kTrancheCap;

    int kept = 0, processed = 0, n_tranche = 0;
    uint64_t hi_incl = ~0ull;
    for (;;) {
        K2_STAMP(0);
        K2_ACC_RESET();
        // ---- 1. select the next tranche ------------------------------------------------------------------
        uint64_t lo = 0;
        if (M - processed > kTrancheCap) lo = select_lower_bound(S, keys, M, smin, hi_incl, nbits);
        K2_STAMP(1);
        if (tid == 0) S.n_sel = 0;
        __syncthreads();
        for (int i0 = 0; i0 < M; i0 += kKeyBatch * kThreads) {
            uint64_t kk[kKeyBatch];
#pragma unroll
            for (int q = 0; q < kKeyBatch; ++q) {
                const int i = i0 + q * kThreads + tid;
                kk[q] = i < M ? __ldg(keys + i) : 0ull;
            }
#pragma unroll
            for (int q = 0; q < kKeyBatch; ++q) {
                const uint64_t nk = norm_key(kk[q], smin);
                const bool in = (i0 + q * kThreads + tid < M) && nk >= lo && nk <= hi_incl;
                const unsigned bal = __ballot_sync(0xffffffffu, in);
                if (bal) {
                    int base = 0;
                    if ((tid & 31) == 0) base = atomicAdd(&S.n_sel, __popc(bal));
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (in) {
                        const int at = base + __popc(bal & ((1u << (tid & 31)) - 1u));
                        if (at < kTrancheCap) S.keys[at] = kk[q];
                    }
                }
            }
        }
        __syncthreads();
        K2_STAMP(2);
        const int n = min(S.n_sel, kTrancheCap);
        n_tranche = n;
        // ---- 2. sort descending, decode boxes -------------------------------------------------------------
        if (n <= kThreads) bitonic_sort_desc<1>(S.keys, n);
        else if (n <= 2 * kThreads) bitonic_sort_desc<2>(S.keys, n);
        else bitonic_sort_desc<4>(S.keys, n);
        const int n_use = min(n, limit - processed);
        K2_STAMP(3);
        int decoded_upto = 0;  // boxes are decoded 1024 at a time, only as far as the NMS walk gets
        // ---- 3. greedy NMS over the tranche, 64 candidates at a time ---------------------------------------
        for (int c0 = 0; c0 < n_use && kept < max_det; c0 += kChunk) {
            const int cn = min(kChunk, n_use - c0);
            K2_ACC_BEGIN();
            if (c0 + cn > decoded_upto) {
                if (!ARRAY && P.family == YSB_YOLOV8 && P.input_kind == YSB_INPUT_RAW_HEADS && P.dfl_bins <= 16) {
                    // DFL decode of exactly this chunk.  A box needs 4 sides x 16 bins = 64 scattered loads: (i) sixteen
                    // threads per candidate fetch 4 bin logits each (all independent, one round trip) into shared memory
                    // laid out [bin][candidate*4 + side]; (ii) one thread per (candidate, side) runs the softmax
                    // expectation from shared memory (conflict-free: consecutive threads, consecutive words) in the
                    // same bin order as the decode kernel; (iii) the four sides of a candidate meet by shuffle.
                    float *dfl = S.area;  // 16 x 256 floats; the areas are only used by the post-filter, after this loop
                    {
                        const int j = tid >> 4, part = tid & 15, side = part >> 2, b0 = (part & 3) << 2;
                        if (j < cn) {
                            int cand = static_cast<int>(key_cand(S.keys[c0 + j]));
                            const Plan &Q = pass_of(P, X, cand);
                            const LevelDesc &lv = Q.lv[find_level(Q, cand)];
                            const float *q = lv.p0 + (static_cast<size_t>(img) * Q.cls_nch + static_cast<size_t>(side) * Q.dfl_bins) * lv.hw +
                                             (cand - lv.cand_off);
                            float v[4];
#pragma unroll
                            for (int b = 0; b < 4; ++b)
                                v[b] = (b0 + b) < Q.dfl_bins ? __ldg(q + static_cast<size_t>(b0 + b) * lv.hw) : -INFINITY;
#pragma unroll
                            for (int b = 0; b < 4; ++b) dfl[(b0 + b) * 256 + j * 4 + side] = v[b];
                        }
                    }
                    __syncthreads();
                    if (tid < 256) {
                        const int j = tid >> 2, side = tid & 3;
                        float sv = 0.0f;
                        if (j < cn) {
                            float v[16];
#pragma unroll
                            for (int i = 0; i < 16; ++i) v[i] = dfl[i * 256 + tid];
                            sv = v8_side_from_bins(v, P.dfl_bins);
                        }
                        const unsigned qb = (tid & 31) & ~3u;
                        const float s0 = __shfl_sync(0xffffffffu, sv, qb), s1 = __shfl_sync(0xffffffffu, sv, qb + 1);
                        const float s2 = __shfl_sync(0xffffffffu, sv, qb + 2), s3 = __shfl_sync(0xffffffffu, sv, qb + 3);
                        if (side == 0 && j < cn) {
                            int cand = static_cast<int>(key_cand(S.keys[c0 + j]));
                            const Plan &Q = pass_of(P, X, cand);
                            S.raw[c0 + j] = tta_undo(Q, v8_box_from_sides(Q, cand, s0, s1, s2, s3));
                        }
                    }
                    decoded_upto = c0 + cn;
                } else {
                    const int i = decoded_upto + tid;
                    if (i < n_use) {
                        int cand = static_cast<int>(key_cand(S.keys[i]));
                        if (ARRAY) {
                            S.raw[i] = __ldg(aa.boxes + cand);
                        } else {
                            const Plan &Q = pass_of(P, X, cand);
                            S.raw[i] = candidate_xyxy(Q, img, cand);
                        }
                    }
                    decoded_upto = min(n_use, decoded_upto + kThreads);
                }
                __syncthreads();
            }
            K2_ACC(8);
            // phase A: 16 threads per candidate test it against the kept list
            {
                const int j = tid >> 4, sub = tid & 15;
                bool sup = false;
                OffBox ob;
                bool valid = j < cn;
                float score = 0.f;
                if (valid) {
                    const uint64_t key = S.keys[c0 + j];
                    score = key_score(key);
                    const float off = P.class_aware ? __fmul_rn(static_cast<float>(key_cls(key)), 4096.0f) : 0.0f;
                    ob = make_offbox(S.raw[c0 + j], off);
                    for (int k = sub; k < kept; k += 16) sup |= pair_hit<ARRAY>(kept_box, k, ob, thr, aa);
                }
                const unsigned bal = __ballot_sync(0xffffffffu, sup);
                const unsigned half = (tid & 16) ? 0xffff0000u : 0x0000ffffu;
                if (sub == 0 && j < kChunk) {
                    S.chunk_pred[j][0] = 0u;
                    S.chunk_pred[j][1] = 0u;
                    // a zero score is never picked by "while sum > 0" (utils/nms.py:16): it is neither kept nor a suppressor
                    S.chunk_alive[j] = valid && !(bal & half) && score > 0.0f;
                    if (valid) soa_store(chunk_box, j, ob);
                }
            }
            __syncthreads();
            K2_ACC(9);
            // phase B: in-chunk suppression bitmask, mask[i] bit j (j > i) = IoU(i, j) reaches the threshold
            {
                const int i = tid >> 4, sub = tid & 15;
                uint32_t lo32 = 0, hi32 = 0;
                if (i < cn && S.chunk_alive[i]) {
                    const OffBox bi = soa_load(chunk_box, i);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int j = q * 16 + sub;  // lanes walk consecutive boxes: conflict-free shared loads
                        if (j > i && j < cn && S.chunk_alive[j] && pair_hit<ARRAY>(chunk_box, j, bi, thr, aa)) {
                            if (j < 32) lo32 |= 1u << j; else hi32 |= 1u << (j - 32);
                            atomicOr(&S.chunk_pred[j][i >> 5], 1u << (i & 31));  // rare: suppression inside a chunk
                        }
                    }
                }
                const unsigned half = (tid & 16) ? 0xffff0000u : 0x0000ffffu;
                lo32 = __reduce_or_sync(half, lo32);
                hi32 = __reduce_or_sync(half, hi32);
                if (sub == 0) S.chunk_mask[i] = (static_cast<uint64_t>(hi32) << 32) | lo32;
            }
            __syncthreads();
            K2_ACC(10);
            // phase C: resolve the chunk (warp 0)
            if (tid < 32) {
                const unsigned a_lo = __ballot_sync(0xffffffffu, S.chunk_alive[tid] != 0);
                const unsigned a_hi = __ballot_sync(0xffffffffu, S.chunk_alive[tid + 32] != 0);
                const uint64_t alive = (static_cast<uint64_t>(a_hi) << 32) | a_lo;
                const bool conflict = (((alive >> tid) & 1ull) && (S.chunk_mask[tid] & alive)) ||
                                      (((alive >> (tid + 32)) & 1ull) && (S.chunk_mask[tid + 32] & alive));
                const unsigned any = __ballot_sync(0xffffffffu, conflict);
                uint64_t keep = alive;
                if (any) {
                    // Greedy resolution without a serial walk: a candidate is decided as soon as all of its earlier
                    // in-chunk suppressors are decided -- kept if none of them was kept, dropped otherwise.  The
                    // lowest undecided candidate is always decidable, so this ends; typical chunks need 2-4 rounds.
                    const uint64_t p0 = (static_cast<uint64_t>(S.chunk_pred[tid][1]) << 32 | S.chunk_pred[tid][0]) & alive;
                    const uint64_t p1 = (static_cast<uint64_t>(S.chunk_pred[tid + 32][1]) << 32 | S.chunk_pred[tid + 32][0]) & alive;
                    uint64_t U = alive, K = 0;
                    while (U) {
                        const bool in0 = (U >> tid) & 1ull, in1 = (U >> (tid + 32)) & 1ull;
                        const bool sup0 = in0 && (p0 & K), sup1 = in1 && (p1 & K);
                        const bool kp0 = in0 && !(p0 & K) && !(p0 & U), kp1 = in1 && !(p1 & K) && !(p1 & U);
                        const uint64_t Rk = (static_cast<uint64_t>(__ballot_sync(0xffffffffu, kp1)) << 32) | __ballot_sync(0xffffffffu, kp0);
                        const uint64_t Rs = (static_cast<uint64_t>(__ballot_sync(0xffffffffu, sup1)) << 32) | __ballot_sync(0xffffffffu, sup0);
                        K |= Rk;
                        U &= ~(Rk | Rs);
                    }
                    keep = K;
                }
                if (tid == 0) {
                    const int room = max_det - kept;
                    int cntk = __popcll(keep);
                    while (cntk > room) {  // drop the lowest-priority (highest index) keeps beyond max_det
                        keep &= ~(1ull << (63 - __clzll(static_cast<long long>(keep))));
                        --cntk;
                    }
                    S.keep_mask = keep;
                }
            }
            __syncthreads();
            K2_ACC(11);
            const uint64_t keep = S.keep_mask;
            if (tid < kChunk && ((keep >> tid) & 1ull)) {
                const int at = kept + __popcll(keep & ((1ull << tid) - 1ull));
                soa_store(kept_box, at, soa_load(chunk_box, tid));
                if (ARRAY) {
                    aa.keep[at] = static_cast<int32_t>(key_cand(S.keys[c0 + tid]));
                } else {
                    S.kept_key[at] = S.keys[c0 + tid];
                    S.kept_raw[at] = S.raw[c0 + tid];
                }
            }
            kept += __popcll(keep);
            __syncthreads();
            K2_ACC(12);
        }
        if (window) {  // the count filter below needs every survivor's box (single tranche holds them all)
            for (int i = decoded_upto + tid; i < n_use; i += kThreads) {
                int cand = static_cast<int>(key_cand(S.keys[i]));
                const Plan &Q = pass_of(P, X, cand);
                S.raw[i] = candidate_xyxy(Q, img, cand);
            }
            __syncthreads();
        }
        K2_STAMP(4);
        K2_ACC_FLUSH();
        processed += n;
        if (kept >= max_det || processed >= limit || processed >= M) break;
        hi_incl = lo - 1;  // lo > 0 here: keys remain below it
    }

    if (ARRAY) {
        if (tid == 0) *aa.keep_cnt = kept;
        return;
    }
    // ---- 5. postprocess_bbox: a kept box survives iff more than one candidate overlaps it by > thr ----------
    if (window) {
        // 1 < M < 3000 <= tranche capacity: the single tranche holds every survivor, boxes decoded for all
        const int warp = tid >> 5, lane = tid & 31;
        const int mm = min(n_tranche, limit);
        if (!P.merge_boxes) {
            // offset boxes + areas once per survivor (in place: the raw boxes of the kept rows live in kept_raw).
            // The 4096-px class offset keeps boxes of different classes disjoint in x as long as the raw boxes span at
            // most 4095 px: for classes a < b, fl(x2 + 4096 a) - fl(x1' + 4096 b) <= (max x2 - min x1) - 4096 + 0.5 < 0
            // (offset values stay below 2^23, so each rounding moves them by at most 0.25).  Then a kept box only needs
            // the survivors of its own class: bucket the survivors by class (counting sort of their indices in shared
            // memory) and walk one bucket per kept box instead of all M survivors.  Otherwise (huge boxes -- the
            // reference then lets classes interact) every pair is tested.
            int lo_x = 0x7fffffff, hi_x = static_cast<int>(0x80000000u);
            bool finite = true;
            for (int j = tid; j < mm; j += kThreads) {
                const uint64_t key = S.keys[j];
                const float off = P.class_aware ? __fmul_rn(static_cast<float>(key_cls(key)), 4096.0f) : 0.0f;
                const float4 rb = S.raw[j];
                finite = finite && (fabsf(rb.x) <= 3.0e38f) && (fabsf(rb.z) <= 3.0e38f);  // false for NaN / inf
                lo_x = min(lo_x, float_ordered(rb.x));
                hi_x = max(hi_x, float_ordered(rb.z));
                const OffBox b = make_offbox(rb, off);
                S.raw[j] = make_float4(b.x1, b.y1, b.x2, b.y2);
                S.area[j] = b.area;
            }
            lo_x = __reduce_min_sync(0xffffffffu, lo_x);
            hi_x = __reduce_max_sync(0xffffffffu, hi_x);
            if (tid == 0) { S.sel_digit = 0x7fffffff; S.sel_above = static_cast<int>(0x80000000u); }
            const bool all_finite = __syncthreads_and(finite ? 1 : 0);
            if (lane == 0) { atomicMin(&S.sel_digit, lo_x); atomicMax(&S.sel_above, hi_x); }
            __syncthreads();
            const float span = __fsub_rn(ordered_float(S.sel_above), ordered_float(S.sel_digit));
            const bool buckets = P.class_aware && thr.positive && P.C <= kThreads && all_finite && span <= 4095.0f;
            if (buckets) {
                uint32_t *cnt_cls = S.hist;                 // [0, C]: bucket starts; [kThreads, kThreads + C): fill cursors
                uint16_t *order = reinterpret_cast<uint16_t *>(S.raw + 3072);  // mm < 3000: the tail of raw[] is free
                cnt_cls[tid] = 0;
                __syncthreads();
                for (int j = tid; j < mm; j += kThreads) atomicAdd(&cnt_cls[key_cls(S.keys[j])], 1u);
                __syncthreads();
                // exclusive scan over the classes (one per thread)
                const uint32_t mine = cnt_cls[tid];
                uint32_t incl = mine;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= d) incl += v;
                }
                if (lane == 31) S.warp_tmp[warp] = incl;
                __syncthreads();
                if (warp == 0) {
                    uint32_t w = S.warp_tmp[lane];
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const uint32_t v = __shfl_up_sync(0xffffffffu, w, d);
                        if (lane >= d) w += v;
                    }
                    S.warp_tmp[lane] = w;  // inclusive over warps
                }
                __syncthreads();
                const uint32_t start = incl - mine + (warp > 0 ? S.warp_tmp[warp - 1] : 0u);
                cnt_cls[tid] = start;
                S.hist[kThreads + tid] = start;
                __syncthreads();
                for (int j = tid; j < mm; j += kThreads) {
                    const uint32_t at = atomicAdd(&S.hist[kThreads + key_cls(S.keys[j])], 1u);
                    order[at] = static_cast<uint16_t>(j);
                }
                __syncthreads();
                for (int r = warp; r < kept; r += kThreads / 32) {
                    const OffBox br = soa_load(kept_box, r);
                    const uint32_t cls_r = key_cls(S.kept_key[r]);
                    const int lo = static_cast<int>(cnt_cls[cls_r]), hi = static_cast<int>(S.hist[kThreads + cls_r]);
                    int c = 0;
                    for (int q = lo + lane; q < hi; q += 32) {
                        const int j = order[q];
                        const float4 o = S.raw[j];
                        const float dw = __fsub_rn(fminf(br.x2, o.z), fmaxf(br.x1, o.x));
                        const float dh = __fsub_rn(fminf(br.y2, o.w), fmaxf(br.y1, o.y));
                        if (!(dw > 0.0f && dh > 0.0f)) continue;
                        c += iou_decide<true>(dw, dh, br.area, S.area[j], thr) ? 1 : 0;
                    }
                    c = __reduce_add_sync(0xffffffffu, c);
                    if (lane == 0) S.kept_flag[r] = c > 1;
                }
            } else {
            for (int r = warp; r < kept; r += kThreads / 32) {
                const OffBox br = soa_load(kept_box, r);
                int c = 0;
                for (int j = lane; j < mm; j += 32) {
                    const float4 o = S.raw[j];  // offset box (x1, y1, x2, y2)
                    const float dw = __fsub_rn(fminf(br.x2, o.z), fmaxf(br.x1, o.x));
                    const float dh = __fsub_rn(fminf(br.y2, o.w), fmaxf(br.y1, o.y));
                    if (thr.positive && !(dw > 0.0f && dh > 0.0f)) continue;
                    c += iou_decide<true>(dw, dh, br.area, S.area[j], thr) ? 1 : 0;
                }
                c = __reduce_add_sync(0xffffffffu, c);
                if (lane == 0) S.kept_flag[r] = c > 1;
            }
            }
        } else {
        for (int r = warp; r < kept; r += kThreads / 32) {
            const OffBox br = soa_load(kept_box, r);
            int c = 0;
            float ax = 0.f, ay = 0.f, az = 0.f, aw = 0.f, ws = 0.f;
            for (int j = lane; j < mm; j += 32) {
                const uint64_t key = S.keys[j];
                const float off = P.class_aware ? __fmul_rn(static_cast<float>(key_cls(key)), 4096.0f) : 0.0f;
                const float4 rj = S.raw[j];
                if (iou_reaches<true>(br, make_offbox(rj, off), thr)) {
                    ++c;
                    // trainer/eval_retinanet.py:346-349 (float32 weights, float32 dot)
                    const float w = key_score(key);
                    ax = __fadd_rn(ax, __fmul_rn(w, rj.x));
                    ay = __fadd_rn(ay, __fmul_rn(w, rj.y));
                    az = __fadd_rn(az, __fmul_rn(w, rj.z));
                    aw = __fadd_rn(aw, __fmul_rn(w, rj.w));
                    ws = __fadd_rn(ws, w);
                }
            }
            c = __reduce_add_sync(0xffffffffu, c);
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                ax = __fadd_rn(ax, __shfl_xor_sync(0xffffffffu, ax, d));
                ay = __fadd_rn(ay, __shfl_xor_sync(0xffffffffu, ay, d));
                az = __fadd_rn(az, __shfl_xor_sync(0xffffffffu, az, d));
                aw = __fadd_rn(aw, __shfl_xor_sync(0xffffffffu, aw, d));
                ws = __fadd_rn(ws, __shfl_xor_sync(0xffffffffu, ws, d));
            }
            if (lane == 0) {
                S.kept_flag[r] = c > 1;
                const float den = __fadd_rn(ws, 1e-16f);
                S.kept_raw[r] = make_float4(__fdiv_rn(ax, den), __fdiv_rn(ay, den), __fdiv_rn(az, den), __fdiv_rn(aw, den));
            }
        }
        }
    } else {
        for (int r = tid; r < kept; r += kThreads) S.kept_flag[r] = 1;
    }
    if (tid == 0) S.out_count = 0;
    __syncthreads();
    K2_STAMP(5);
    // ---- ordered write of the surviving rows ---------------------------------------------------------------
    {
        const int r = tid;  // kept <= kMaxKeep == kThreads
        bool pass = false;
        float4 box = make_float4(0.f, 0.f, 0.f, 0.f);
        uint64_t key = 0;
        if (r < kept) {
            box = S.kept_raw[r];
            key = S.kept_key[r];
            pass = S.kept_flag[r] != 0;
            if (pass && P.small_box_filter)  // remove_small_boxes, trainer/eval_yolov7.py:203-213
                pass = (__fsub_rn(box.z, box.x) > P.min_box_wh) && (__fsub_rn(box.w, box.y) > P.min_box_wh);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, pass);
        if ((tid & 31) == 0) S.warp_tmp[tid >> 5] = __popc(bal);
        __syncthreads();
        int before = 0;
        for (int w = 0; w < (tid >> 5); ++w) before += S.warp_tmp[w];
        if (pass) {
            const int at = before + __popc(bal & ((1u << (tid & 31)) - 1u));
            const float sc = key_score(key);
            const float s_out = P.topk_sqrt ? sqrtf(sc) : sc;  // FCOS: x[:, 4] = sqrt(x[:, 4]) (trainer/eval_fcos.py:281)
            const float c_out = static_cast<float>(key_cls(key));
            const size_t off = (static_cast<size_t>(img) * max_det + at) * 6;
            // one destination (the caller's buffer) or, in a detection gather, every rank's receive slot: plain stores,
            // local or over NVLink; a warp writes 32 consecutive rows = 768 contiguous bytes
            const int ndst = G.world > 0 ? G.world : 1;
            for (int d = 0; d < ndst; ++d) {
                float *row = (G.world > 0 ? G.rows[d] : dets) + off;
                row[0] = box.x; row[1] = box.y; row[2] = box.z; row[3] = box.w;
                row[4] = s_out;
                row[5] = c_out;
            }
            if (det_idx) det_idx[static_cast<size_t>(img) * max_det + at] = static_cast<int32_t>(key_cand(key));
        }
        K2_STAMP(6);
        if (tid == kThreads - 1) {
            const int total = before + __popc(bal);
            const int value = (total == 0 && P.none_when_empty) ? -1 : total;
            if (G.world > 0) {
                for (int d = 0; d < G.world; ++d) G.cnt[d][img] = value;
            } else {
                det_cnt[img] = value;
            }
        }
        if (G.world > 0) {
            // release: every row / count store of this CTA is ordered before the arrival count each peer will see
            __threadfence_system();
            __syncthreads();
            if (tid < G.world) atomicAdd_system(G.arrived[tid], 1u);
        }
    }
}

cudaError_t launch_select_nms(const Plan &P, const uint64_t *d_keys, int64_t key_cap, const int32_t *d_counts,
                              float *d_dets, int32_t *d_det_idx, int32_t *d_det_cnt, cudaStream_t stream,
                              const GatherSink *sink)
{
    if (P.batch == 0) return cudaSuccess;
    // per-device attribute; set on every launch (host-side, sub-microsecond) so that a process driving several
    // devices never launches with the default 48 KB limit
    cudaError_t e = cudaFuncSetAttribute(k_select_nms<false, NoExtraPasses>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(sizeof(NmsSmem)));
    if (e != cudaSuccess) return e;
    ArrayArgs aa{};
    GatherSink G;
    if (sink) G = *sink;
    else memset(&G, 0, sizeof(G));
    k_select_nms<false, NoExtraPasses><<<P.batch, kThreads, sizeof(NmsSmem), stream>>>(
        P, NoExtraPasses{}, d_keys, key_cap, d_counts, d_dets, d_det_idx, d_det_cnt, aa, G);
    return cudaGetLastError();
}

// test-time augmentation: one selection/NMS over the key lists the passes appended to; X describes passes 1..n
cudaError_t launch_select_nms_tta(const Plan &P, const ExtraPasses &X, const uint64_t *d_keys, int64_t key_cap,
                                  const int32_t *d_counts, float *d_dets, int32_t *d_det_idx, int32_t *d_det_cnt,
                                  cudaStream_t stream)
{
    if (P.batch == 0) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(k_select_nms<false, ExtraPasses>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(sizeof(NmsSmem)));
    if (e != cudaSuccess) return e;
    ArrayArgs aa{};
    GatherSink G;
    memset(&G, 0, sizeof(G));
    k_select_nms<false, ExtraPasses><<<P.batch, kThreads, sizeof(NmsSmem), stream>>>(P, X, d_keys, key_cap, d_counts, d_dets,
                                                                                      d_det_idx, d_det_cnt, aa, G);
    return cudaGetLastError();
}

// ---- array flavour: utils.numba_nms / utils.gpu_nms over one explicit (m,4)/(m) pair ---------------------------
// zero scores are never kept (utils/nms.py:16 "while score.sum() > 0"); keys carry the array index.
__global__ void k_array_keys(const float *__restrict__ scores, int m, uint64_t *__restrict__ keys, int32_t *__restrict__ counts)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool ok = i < m && scores[i] > 0.0f;
    const unsigned bal = __ballot_sync(0xffffffffu, ok);
    if (!bal) return;
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t sb = ok ? __float_as_uint(scores[i]) : 0u;
    const uint32_t wmax = __reduce_max_sync(0xffffffffu, sb);
    const uint32_t wmin = __reduce_max_sync(0xffffffffu, ok ? ~sb : 0u);
    int base = 0;
    if (lane == 0) {
        base = atomicAdd(counts, __popc(bal));
        atomicMax(reinterpret_cast<unsigned int *>(counts + 2), wmax);
        atomicMax(reinterpret_cast<unsigned int *>(counts + 3), wmin);
    }
    base = __shfl_sync(0xffffffffu, base, 0);
    if (ok) keys[base + __popc(bal & ((1u << lane) - 1u))] = pack_key(scores[i], static_cast<uint32_t>(i), 0u);
}

static size_t align256(size_t x) { return (x + 255) & ~static_cast<size_t>(255); }

size_t array_nms_workspace_bytes(int64_t m)
{
    const size_t mm = static_cast<size_t>(m > 0 ? m : 1);
    return align256(sizeof(uint64_t) * mm) + 3 * align256(sizeof(float2) * mm) + align256(sizeof(int32_t) * 4) + 256;
}

cudaError_t launch_array_nms(const float *d_boxes, const float *d_scores, int64_t m, double iou_thr, int cmp, int iou_kind,
                             int64_t max_keep, void *ws, int32_t *d_keep, int32_t *d_keep_cnt, cudaStream_t stream)
{
    if (m == 0) return cudaMemsetAsync(d_keep_cnt, 0, sizeof(int32_t), stream);
    const size_t mm = static_cast<size_t>(m);
    uintptr_t base = (reinterpret_cast<uintptr_t>(ws) + 255) & ~static_cast<uintptr_t>(255);
    uint64_t *keys = reinterpret_cast<uint64_t *>(base);
    const size_t seg = align256(sizeof(float2) * mm);
    uintptr_t kb = base + align256(sizeof(uint64_t) * mm);
    BoxSoA kept{reinterpret_cast<float2 *>(kb), reinterpret_cast<float2 *>(kb + seg), reinterpret_cast<float *>(kb + 2 * seg)};
    int32_t *counts = reinterpret_cast<int32_t *>(kb + 3 * seg);
    cudaError_t e = cudaMemsetAsync(counts, 0, sizeof(int32_t) * 4, stream);
    if (e != cudaSuccess) return e;
    k_array_keys<<<static_cast<unsigned>((m + 255) / 256), 256, 0, stream>>>(d_scores, static_cast<int>(m), keys, counts);
    e = cudaFuncSetAttribute(k_select_nms<true, NoExtraPasses>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             static_cast<int>(sizeof(NmsSmem)));
    if (e != cudaSuccess) return e;
    Plan P;
    memset(&P, 0, sizeof(P));
    P.batch = 1;
    P.N = static_cast<int>(m);
    P.iou_thr = iou_thr;
    P.max_det = static_cast<int>((max_keep > 0 && max_keep < m) ? max_keep : m);
    ArrayArgs aa;
    aa.boxes = reinterpret_cast<const float4 *>(d_boxes);
    aa.kept = kept;
    aa.keep = d_keep;
    aa.keep_cnt = d_keep_cnt;
    aa.iou_kind = iou_kind;
    aa.cmp = cmp;
    aa.thr32 = static_cast<float>(iou_thr);
    GatherSink G;
    memset(&G, 0, sizeof(G));
    k_select_nms<true, NoExtraPasses><<<1, kThreads, sizeof(NmsSmem), stream>>>(P, NoExtraPasses{}, keys, m, counts, nullptr,
                                                                                 nullptr, nullptr, aa, G);
    return cudaGetLastError();
}

#ifdef YSB_K2_TIMING
cudaError_t debug_k2_timing(long long *host_out) { return cudaMemcpyFromSymbol(host_out, g_k2_timing, sizeof(g_k2_timing)); }
#endif

}  // namespace ysb

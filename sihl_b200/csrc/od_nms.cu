// od_nms.cu — class-aware greedy NMS (north-star extension, the
// reference has no NMS).  Semantics: torchvision's exact per-class path
// (_batched_nms_vanilla, tv:ops/boxes.py:102-120, over the CPU nms kernel): visit boxes by
// (score desc, index asc); a kept box suppresses every later box of the same class with
// inter / (area_i + area_j - inter) > thr.
//
// One CTA per image (segment), everything for that image on chip when it fits:
//   1. bitonic sort of (score-key | ~index, slot) descending         -> rank r (output order)
//   2. bitonic sort of (class << 32 | r) ascending                   -> class segments, score
//      order preserved inside each (the rank makes the second sort stable)
//   3. greedy suppression inside each class segment only: one warp per segment, the kept box
//      is broadcast and 32 later boxes are tested per step (dead flags in shared memory);
//      segments longer than kBigSegment are swept by the whole CTA instead
//   4. ordered compaction of the survivors by rank r               -> keep list / top-K rows
// Lists longer than kSmemItems spill the same arrays to a global workspace (same code,
// generic pointers).  Work is O(sum_c n_c^2 / 32) warp steps instead of the O(N^2) pair
// tests of the all-pairs bitmask formulation.
#include "od_common.cuh"

namespace sihl {

// Developer instrumentation (compiled out unless -DSIHL_PHASE_TIMING): SM clock of block 0 /
// thread 0 at phase boundaries, read back with sihl_od_debug_phases().
#ifdef SIHL_PHASE_TIMING
__device__ long long g_phase_clock[16];
#define SIHL_PHASE(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_phase_clock[i] = clock64(); } while (0)
#else
#define SIHL_PHASE(i) do { } while (0)
#endif

constexpr int kNmsThreads = 1024;
constexpr int kSmemItems = 4096;          // per-image candidates held on chip
constexpr int kBigSegment = 512;
constexpr int kSmallItems = 256;          // lists up to this size take the single-pass rank-sort kernel
static_assert(4 * kSmallItems == kNmsThreads, "short-list path: four lanes per candidate");
constexpr size_t kSmemItemBytes = 41;     // key 8 + box 16 + val 4 + area 4 + slot 4 + seg 4 + dead 1
constexpr size_t kWsItemBytes = 48;       // workspace stride per item (keeps every image 16-B aligned)

__device__ __forceinline__ unsigned f2ord_nms(float f)
{
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

struct NmsArrays {
    unsigned long long *key;   // [np] sort keys; after step 2: (class << 32 | r) by q
    float4 *box;               // [np] by q
    unsigned *val;             // [np] rank -> slot, later rank -> q
    float *area;               // [np] by q
    unsigned *slot;            // [np] by q
    int *seg;                  // [np] segment starts (+ sentinel)
    unsigned char *dead;       // [np] by q
};

__device__ __forceinline__ NmsArrays carve(unsigned char *base, int np)
{
    NmsArrays a;
    a.key = reinterpret_cast<unsigned long long *>(base);
    a.box = reinterpret_cast<float4 *>(base + (size_t)np * 8);
    a.val = reinterpret_cast<unsigned *>(base + (size_t)np * 24);
    a.area = reinterpret_cast<float *>(base + (size_t)np * 28);
    a.slot = reinterpret_cast<unsigned *>(base + (size_t)np * 32);
    a.seg = reinterpret_cast<int *>(base + (size_t)np * 36);
    a.dead = base + (size_t)np * 40;
    return a;
}

template <bool DESC>
__device__ __forceinline__ void bitonic_kv(unsigned long long *key, unsigned *val, int np)
{
    for (int k = 2; k <= np; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < np; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long a = key[i], b = key[ixj];
                    const bool first = (i & k) == 0;
                    const bool swap = DESC ? (first ? (a < b) : (a > b)) : (first ? (a > b) : (a < b));
                    if (swap) {
                        key[i] = b; key[ixj] = a;
                        if (val) { const unsigned t = val[i]; val[i] = val[ixj]; val[ixj] = t; }
                    }
                }
            }
            __syncthreads();
        }
}

// Ordered block compaction: emit(i, rank) for every i in [0,n) with flag(i); rank ascends with i.
template <class Flag, class Emit>
__device__ __forceinline__ int block_ordered_compact(int n, int *s_warp /* [33] */, Flag flag, Emit emit)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    int running = 0;
    for (int i0 = 0; i0 < n; i0 += blockDim.x) {
        const int i = i0 + threadIdx.x;
        const bool f = i < n && flag(i);
        const unsigned m = __ballot_sync(kFullMask, f);
        if (lane == 0) s_warp[warp] = __popc(m);
        __syncthreads();
        if (warp == 0) {
            const int x = lane < nwarps ? s_warp[lane] : 0;
            int incl = x;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(kFullMask, incl, o);
                if (lane >= o) incl += y;
            }
            if (lane < nwarps) s_warp[lane] = incl - x;
            if (lane == 31) s_warp[32] = incl;
        }
        __syncthreads();
        if (f) emit(i, running + s_warp[warp] + __popc(m & ((1u << lane) - 1u)));
        running += s_warp[32];
        __syncthreads();
    }
    return running;
}

// torchvision/csrc/ops/cpu/nms_kernel.cpp: ovr = inter / (iarea + areas[j] - inter); ovr > thr
__device__ __forceinline__ bool iou_gt(float4 a, float area_a, float4 b, float area_b, float thr)
{
    const float w = fmaxf(0.f, fminf(a.z, b.z) - fmaxf(a.x, b.x));
    const float h = fmaxf(0.f, fminf(a.w, b.w) - fmaxf(a.y, b.y));
    const float inter = w * h;
    return inter / ((area_a + area_b) - inter) > thr;
}

struct NmsParams {
    // mode 0: candidate lists written by k_dense_decode
    const int32_t *cand_count; int64_t cap;
    const unsigned long long *cand_key; const float4 *cand_box; const int32_t *cand_cls;
    // mode 1: torchvision.ops.batched_nms inputs + segment offsets
    const float4 *boxes; const float *scores; const int64_t *classes; const int32_t *seg_offsets;
    int mode;
    float iou_thr;
    int K; int64_t *num_instances; float *out_scores; int64_t *out_classes; float4 *out_boxes;   // mode 0
    int64_t *keep; int32_t *keep_count;                                                           // mode 1
    unsigned char *workspace; size_t ws_stride;   // mode 0: bytes per image; mode 1: unused (offset = 2*seg start)
    int reset_counts;          // mode 0: zero cand_count[img] once consumed (saves the next step's memset launch)
};

__device__ __forceinline__ void nms_big_body(const NmsParams &p)
{
    extern __shared__ __align__(16) unsigned char s_dyn[];
    __shared__ int s_warp[33];

    const int img = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    int n, src0;
    if (p.mode == 0) {
        const int c = __ldg(p.cand_count + img);
        n = (int)(c < p.cap ? c : p.cap);
        src0 = 0;
    } else {
        src0 = __ldg(p.seg_offsets + img);
        n = __ldg(p.seg_offsets + img + 1) - src0;
    }
    int np = 2;
    while (np < n) np <<= 1;
    int n_kept = 0;

    if (n > 0) {
        unsigned char *base = s_dyn;
        if (np > kSmemItems)
            base = p.workspace + (p.mode == 0 ? (size_t)img * p.ws_stride : (size_t)2 * src0 * kWsItemBytes);
        const NmsArrays ar = carve(base, np);

        // 1. rank by (score desc, index asc)
        for (int i = tid; i < np; i += blockDim.x) {
            unsigned long long key = 0ull;
            if (i < n) {
                if (p.mode == 0) key = __ldg(p.cand_key + (int64_t)img * p.cap + i);
                else key = ((unsigned long long)f2ord_nms(__ldg(p.scores + src0 + i)) << 32) |
                           (unsigned long long)(0xffffffffu - (unsigned)i);
            }
            ar.key[i] = key;
            ar.val[i] = (unsigned)i;
        }
        __syncthreads();
        bitonic_kv<true>(ar.key, ar.val, np);

        // 2. class-major order; the rank in the low word keeps score order inside a class
        for (int r = tid; r < np; r += blockDim.x) {
            unsigned long long key = ~0ull;
            if (r < n) {
                const unsigned slot = ar.val[r];
                const unsigned c = p.mode == 0 ? (unsigned)__ldg(p.cand_cls + (int64_t)img * p.cap + slot)
                                               : (unsigned)__ldg(p.classes + src0 + slot);
                key = ((unsigned long long)c << 32) | (unsigned long long)r;
            }
            ar.key[r] = key;
        }
        __syncthreads();
        bitonic_kv<false>(ar.key, nullptr, np);
        for (int q = tid; q < n; q += blockDim.x) {
            const unsigned r = (unsigned)(ar.key[q] & 0xffffffffu), slot = ar.val[r];
            const float4 bx = p.mode == 0 ? __ldg(p.cand_box + (int64_t)img * p.cap + slot) : __ldg(p.boxes + src0 + slot);
            ar.box[q] = bx;
            ar.area[q] = (bx.z - bx.x) * (bx.w - bx.y);
            ar.slot[q] = slot;
            ar.dead[q] = 0;
        }
        __syncthreads();
        for (int q = tid; q < n; q += blockDim.x) ar.val[(unsigned)(ar.key[q] & 0xffffffffu)] = (unsigned)q;   // rank -> q

        // 3. class segments
        const int n_seg = block_ordered_compact(
            n, s_warp, [&](int q) { return q == 0 || (ar.key[q] >> 32) != (ar.key[q - 1] >> 32); },
            [&](int q, int s) { ar.seg[s] = q; });
        if (tid == 0) ar.seg[n_seg] = n;
        __syncthreads();
        bool any_big = false;
        for (int s = warp; s < n_seg; s += nwarps) {
            const int q0 = ar.seg[s], m = ar.seg[s + 1] - q0;
            if (m > kBigSegment) continue;
            for (int i = 0; i + 1 < m; ++i) {
                __syncwarp();
                if (ar.dead[q0 + i]) continue;
                const float4 bi = ar.box[q0 + i];
                const float ai = ar.area[q0 + i];
                for (int j = i + 1 + lane; j < m; j += 32)
                    if (!ar.dead[q0 + j] && iou_gt(bi, ai, ar.box[q0 + j], ar.area[q0 + j], p.iou_thr)) ar.dead[q0 + j] = 1;
            }
        }
        for (int s = 0; s < n_seg; ++s) any_big |= (ar.seg[s + 1] - ar.seg[s]) > kBigSegment;
        __syncthreads();
        if (any_big) {
            for (int s = 0; s < n_seg; ++s) {
                const int q0 = ar.seg[s], m = ar.seg[s + 1] - q0;
                if (m <= kBigSegment) continue;
                for (int i = 0; i + 1 < m; ++i) {
                    __syncthreads();
                    if (ar.dead[q0 + i]) continue;
                    const float4 bi = ar.box[q0 + i];
                    const float ai = ar.area[q0 + i];
                    for (int j = i + 1 + tid; j < m; j += blockDim.x)
                        if (!ar.dead[q0 + j] && iou_gt(bi, ai, ar.box[q0 + j], ar.area[q0 + j], p.iou_thr))
                            ar.dead[q0 + j] = 1;
                }
                __syncthreads();
            }
        }

        // 4. survivors in rank order
        n_kept = block_ordered_compact(
            n, s_warp, [&](int r) { return ar.dead[ar.val[r]] == 0; },
            [&](int r, int k) {
                const unsigned q = ar.val[r], slot = ar.slot[q];
                if (p.mode == 0) {
                    if (k < p.K) {
                        const int64_t o = (int64_t)img * p.K + k;
                        p.out_scores[o] = __uint_as_float((unsigned)(__ldg(p.cand_key + (int64_t)img * p.cap + slot) >> 32));
                        p.out_classes[o] = (int64_t)(ar.key[q] >> 32);
                        p.out_boxes[o] = ar.box[q];
                    }
                } else {
                    p.keep[src0 + k] = (int64_t)src0 + slot;
                }
            });
    }

    if (p.mode == 0) {
        const int m = n_kept < p.K ? n_kept : p.K;
        if (tid == 0) {
            p.num_instances[img] = m;
            if (p.reset_counts) const_cast<int32_t *>(p.cand_count)[img] = 0;
        }
        for (int k = m + tid; k < p.K; k += blockDim.x) {
            const int64_t o = (int64_t)img * p.K + k;
            p.out_scores[o] = 0.f;
            p.out_classes[o] = 0;
            p.out_boxes[o] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    } else if (tid == 0) {
        p.keep_count[img] = n_kept;
    }
}

// Short lists (n <= kSmallItems, the common case after a 0.05 score threshold).  Four lanes per
// candidate; ONE counting pass over shared memory yields, per candidate, its rank r in
// (score desc, index asc) order, its class-major position q, and the start / length of its
// class segment.  Suppression is then a per-candidate bitmask of earlier same-class boxes with
// IoU > thr (all pairs in parallel, no serial dependence), resolved by one lane per segment with
// pure bit operations — the greedy chain costs a few cycles per box instead of a shared-memory
// round trip per kept box.  Segments longer than 32 fall back to the broadcast loop of k_nms.
__device__ __forceinline__ void nms_small_body(const NmsParams &p)
{
    __shared__ unsigned long long s_key[kSmallItems];
    __shared__ unsigned s_cls[kSmallItems];        // by slot
    __shared__ float4 s_box[kSmallItems];          // by q
    __shared__ float s_area[kSmallItems];          // by q
    __shared__ unsigned s_mask[kSmallItems];       // by q: earlier boxes of the segment that suppress q (if kept)
    __shared__ unsigned s_keep[kSmallItems];       // by segment start: kept bits of the segment
    __shared__ unsigned char s_dead[kSmallItems];  // by q (only for segments longer than 32)
    __shared__ unsigned short s_q_of_r[kSmallItems], s_slot_of_r[kSmallItems], s_seg0_of_q[kSmallItems], s_len_of_q[kSmallItems];
    __shared__ int s_warp[33];
    __shared__ int s_big;

    const int img = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    const int item = tid >> 2, sub = tid & 3;
    SIHL_PHASE(0);
    int n, src0;
    if (p.mode == 0) {
        const int c = __ldg(p.cand_count + img);
        n = (int)(c < p.cap ? c : p.cap);
        src0 = 0;
    } else {
        src0 = __ldg(p.seg_offsets + img);
        n = __ldg(p.seg_offsets + img + 1) - src0;
    }
    SIHL_PHASE(1);
    int n_kept = 0;
    if (n > 0) {
        const bool have = item < n;
        unsigned long long key = 0ull;
        unsigned cls = 0;
        float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
        if (have) {
            if (p.mode == 0) {
                key = __ldg(p.cand_key + (int64_t)img * p.cap + item);
                cls = (unsigned)__ldg(p.cand_cls + (int64_t)img * p.cap + item);
                bx = __ldg(p.cand_box + (int64_t)img * p.cap + item);
            } else {
                key = ((unsigned long long)f2ord_nms(__ldg(p.scores + src0 + item)) << 32) |
                      (unsigned long long)(0xffffffffu - (unsigned)item);
                cls = (unsigned)__ldg(p.classes + src0 + item);
                bx = __ldg(p.boxes + src0 + item);
            }
            if (sub == 0) { s_key[item] = key; s_cls[item] = cls; }
        }
        if (tid == 0) s_big = 0;
        __syncthreads();
        SIHL_PHASE(2);
        // one pass: r = #better keys; seg0 = #smaller classes; k = #better keys of the same class; len = #same class
        int r = 0, seg0 = 0, k = 0, len = 0;
        if (have) {
#pragma unroll 4
            for (int j = sub; j < n; j += 4) {
                const unsigned long long kj = s_key[j];
                const unsigned cj = s_cls[j];
                const bool better = kj > key, same = cj == cls;
                r += better;
                seg0 += cj < cls;
                k += same && better;
                len += same;
            }
        }
#pragma unroll
        for (int o = 2; o > 0; o >>= 1) {
            r += __shfl_xor_sync(kFullMask, r, o);
            seg0 += __shfl_xor_sync(kFullMask, seg0, o);
            k += __shfl_xor_sync(kFullMask, k, o);
            len += __shfl_xor_sync(kFullMask, len, o);
        }
        const int q = seg0 + k;
        if (have && sub == 0) {
            s_box[q] = bx;
            s_area[q] = (bx.z - bx.x) * (bx.w - bx.y);
            s_q_of_r[r] = (unsigned short)q;
            s_slot_of_r[r] = (unsigned short)item;
            s_seg0_of_q[q] = (unsigned short)seg0;
            s_len_of_q[q] = (unsigned short)len;
            s_dead[q] = 0;
            if (len > 32) s_big = 1;
        }
        __syncthreads();
        SIHL_PHASE(3);
        // suppressor masks: the 4 lanes of a candidate split its earlier same-class boxes
        unsigned mask = 0;
        if (have && len <= 32) {
            const float area = (bx.z - bx.x) * (bx.w - bx.y);
            for (int j = sub; j < k; j += 4)
                if (iou_gt(s_box[seg0 + j], s_area[seg0 + j], bx, area, p.iou_thr)) mask |= 1u << j;
        }
        mask |= __shfl_xor_sync(kFullMask, mask, 1);
        mask |= __shfl_xor_sync(kFullMask, mask, 2);
        if (have && sub == 0) s_mask[q] = mask;
        __syncthreads();
        SIHL_PHASE(4);
        if (have && sub == 0 && k == 0 && len <= 32) {          // segment leader: greedy pass on bits only
            unsigned kept = 1u;
            for (int i = 1; i < len; ++i)
                if ((s_mask[seg0 + i] & kept) == 0u) kept |= 1u << i;
            s_keep[seg0] = kept;
        }
        __syncthreads();
        if (s_big) {                                             // rare: a class with more than 32 candidates
            const int warp = tid >> 5, nwarps = blockDim.x >> 5;
            for (int qq = warp; qq < n; qq += nwarps) {          // a warp per long segment, found by its leader
                const int q0 = s_seg0_of_q[qq], m = s_len_of_q[qq];
                if (qq != q0 || m <= 32) continue;
                for (int i = 0; i + 1 < m; ++i) {
                    __syncwarp();
                    if (s_dead[q0 + i]) continue;
                    const float4 bi = s_box[q0 + i];
                    const float ai = s_area[q0 + i];
                    for (int j = i + 1 + lane; j < m; j += 32)
                        if (!s_dead[q0 + j] && iou_gt(bi, ai, s_box[q0 + j], s_area[q0 + j], p.iou_thr)) s_dead[q0 + j] = 1;
                }
            }
            __syncthreads();
        }
        SIHL_PHASE(5);
        // survivors in rank order
        n_kept = block_ordered_compact(
            n, s_warp,
            [&](int rr) {
                const int qq = s_q_of_r[rr], q0 = s_seg0_of_q[qq];
                return s_len_of_q[qq] <= 32 ? ((s_keep[q0] >> (qq - q0)) & 1u) != 0u : s_dead[qq] == 0;
            },
            [&](int rr, int kk) {
                const int qq = s_q_of_r[rr], slot = s_slot_of_r[rr];
                if (p.mode == 0) {
                    if (kk < p.K) {
                        const int64_t o = (int64_t)img * p.K + kk;
                        p.out_scores[o] = __uint_as_float((unsigned)(s_key[slot] >> 32));
                        p.out_classes[o] = (int64_t)s_cls[slot];
                        p.out_boxes[o] = s_box[qq];
                    }
                } else {
                    p.keep[src0 + kk] = (int64_t)src0 + slot;
                }
            });
    }
    SIHL_PHASE(6);
    if (p.mode == 0) {
        const int m = n_kept < p.K ? n_kept : p.K;
        if (tid == 0) {
            p.num_instances[img] = m;
            if (p.reset_counts) const_cast<int32_t *>(p.cand_count)[img] = 0;
        }
        for (int kk = m + tid; kk < p.K; kk += blockDim.x) {
            const int64_t o = (int64_t)img * p.K + kk;
            p.out_scores[o] = 0.f;
            p.out_classes[o] = 0;
            p.out_boxes[o] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    } else if (tid == 0) {
        p.keep_count[img] = n_kept;
    }
    SIHL_PHASE(7);
}

// One launch for both paths: the list length (known only on the device) picks the body.
__global__ void __launch_bounds__(kNmsThreads) k_nms(NmsParams p)
{
    int n;
    if (p.mode == 0) {
        const int c = __ldg(p.cand_count + blockIdx.x);
        n = (int)(c < p.cap ? c : p.cap);
    } else {
        n = __ldg(p.seg_offsets + blockIdx.x + 1) - __ldg(p.seg_offsets + blockIdx.x);
    }
    if (n <= kSmallItems) nms_small_body(p);
    else nms_big_body(p);
}

static int pow2ceil(int64_t n)
{
    int p = 2;
    while (p < n) p <<= 1;
    return p;
}

static int launch_nms(const NmsParams &p, int n_images, int64_t max_items, cudaStream_t st)
{
    // dynamic shared memory only when a list can exceed the short-list path
    const size_t smem = max_items > kSmallItems ? (size_t)kSmemItems * kSmemItemBytes : 0;
    static thread_local bool attr_set = false;
    if (smem && !attr_set) {
        int rc = cuda_status(cudaFuncSetAttribute(k_nms, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                  (int)((size_t)kSmemItems * kSmemItemBytes)),
                             "cudaFuncSetAttribute(k_nms)");
        if (rc) return rc;
        attr_set = true;
    }
    k_nms<<<n_images, kNmsThreads, smem, st>>>(p);
    SIHL_CHECK_LAUNCH("k_nms");
    return SIHL_OD_OK;
}

}  // namespace sihl

using namespace sihl;

extern "C" size_t sihl_od_nms_workspace_bytes(int batch, int64_t cand_capacity)
{
    if (batch <= 0 || cand_capacity <= 0) return 0;
    const int np = pow2ceil(cand_capacity);
    if (np <= kSmemItems) return 0;
    return (size_t)batch * (size_t)np * kWsItemBytes;
}

extern "C" int sihl_od_nms_topk(const int32_t *cand_count, int64_t cand_capacity, const uint64_t *cand_key,
                                const float *cand_box, const int32_t *cand_cls, int batch, float iou_thr, int k,
                                int64_t *num_instances, float *scores, int64_t *classes, float *boxes, void *workspace,
                                int reset_counts, void *stream)
{
    SIHL_CHECK_ARG(cand_count && cand_key && cand_box && cand_cls, "NULL input");
    SIHL_CHECK_ARG(num_instances && scores && classes && boxes, "NULL output");
    SIHL_CHECK_ARG(cand_capacity >= 1 && cand_capacity < (1ll << 30) && k >= 1, "bad sizes");
    SIHL_CHECK_ARG(workspace != nullptr || sihl_od_nms_workspace_bytes(batch, cand_capacity) == 0,
                   "workspace needed for capacity %lld", (long long)cand_capacity);
    if (batch <= 0) return SIHL_OD_OK;
    NmsParams p = {};
    p.cand_count = cand_count; p.cap = cand_capacity;
    p.cand_key = reinterpret_cast<const unsigned long long *>(cand_key);
    p.cand_box = reinterpret_cast<const float4 *>(cand_box); p.cand_cls = cand_cls;
    p.mode = 0; p.iou_thr = iou_thr; p.K = k; p.reset_counts = reset_counts;
    p.num_instances = num_instances; p.out_scores = scores; p.out_classes = classes;
    p.out_boxes = reinterpret_cast<float4 *>(boxes);
    p.workspace = static_cast<unsigned char *>(workspace);
    p.ws_stride = (size_t)pow2ceil(cand_capacity) * kWsItemBytes;
    return launch_nms(p, batch, cand_capacity, (cudaStream_t)stream);
}

extern "C" size_t sihl_od_batched_nms_workspace_bytes(int64_t n)
{
    if (n <= kSmemItems) return 0;
    return (size_t)(2 * n + 2) * kWsItemBytes;
}

extern "C" int sihl_od_batched_nms(const float *boxes, const float *scores, const int64_t *classes,
                                   const int32_t *seg_offsets, int n_images, int64_t n, float iou_thr, int64_t *keep,
                                   int32_t *keep_count, void *workspace, void *stream)
{
    SIHL_CHECK_ARG(seg_offsets && keep_count, "NULL argument");
    SIHL_CHECK_ARG(n >= 0 && n < (1ll << 30) && n_images >= 0, "bad sizes");
    SIHL_CHECK_ARG(n == 0 || (boxes && scores && classes && keep), "NULL argument");
    SIHL_CHECK_ARG(workspace != nullptr || sihl_od_batched_nms_workspace_bytes(n) == 0, "workspace needed for n=%lld",
                   (long long)n);
    if (n_images == 0) return SIHL_OD_OK;
    NmsParams p = {};
    p.boxes = reinterpret_cast<const float4 *>(boxes); p.scores = scores; p.classes = classes; p.seg_offsets = seg_offsets;
    p.mode = 1; p.iou_thr = iou_thr; p.keep = keep; p.keep_count = keep_count;
    p.workspace = static_cast<unsigned char *>(workspace); p.ws_stride = 0;
    return launch_nms(p, n_images, n, (cudaStream_t)stream);
}

#ifdef SIHL_PHASE_TIMING
extern "C" __attribute__((visibility("default"))) int sihl_od_debug_phases(long long *out_host)
{
    return cudaMemcpyFromSymbol(out_host, sihl::g_phase_clock, sizeof(long long) * 16) == cudaSuccess ? 0 : 2;
}
#endif

// Row N4: the step BEFORE the hot path — the head's per-location MLP towers
// (ref src/sihl/heads/object_detection.py:51-61 builds four torchvision ``ops.MLP(256 -> [256]*4 + [out],
// norm_layer=LayerNorm, activation_layer=SiLU)``; :116 / :121 / :175 apply them to every location).  One tower layer is
//
//     hidden:  y = SiLU(LayerNorm(x W^T + b) * gamma + beta)     x [M,256] bf16, W [256,256] bf16, y [M,256] bf16
//     output:  y = x W^T + b                                     W [n,256] bf16 (n = 1, 4, num_classes), y [M,n] fp32
//
// and this is the one GEMM-shaped piece of the head, so it is the one place the 5th-generation tensor cores are used:
//
//   * W stays resident in shared memory for the life of the (persistent, one-per-SM) CTA — four 64-column K chunks in
//     the 128-byte-swizzled K-major layout `tcgen05.mma` reads through a shared-memory descriptor;
//   * x row tiles (128 rows x 64 columns = 16 KB) stream through a TMA ring (`cp.async.bulk.tensor.2d`, hardware
//     swizzle, mbarrier completion); rows past M are zero-filled by the TMA unit;
//   * one elected thread issues 16 `tcgen05.mma.cta_group::1.kind::f16` (M=128, N, K=16) per tile into one of two
//     TMEM accumulator stages (2 x N fp32 columns), `tcgen05.commit` releases ring slots / publishes the accumulator;
//   * four epilogue warps — one TMEM lane quarter each, thread = row — read the accumulator with `tcgen05.ld`,
//     add the bias, normalise the row (two-pass mean / variance, fp32, entirely inside the thread: no shuffles, no
//     shared memory), apply SiLU and write bf16 rows; the other accumulator stage is being filled meanwhile.
//
// A hidden layer reads 512 B and writes 512 B per location for 131 kFLOP: at M = 545 600 (640^2, batch 64) that is
// 559 MB and 71.5 GFLOP per layer — HBM-bound at ~87 us, a third of what the unfused Linear + LayerNorm + SiLU kernels
// move.  Every wait is bounded: a pipeline bug traps instead of hanging the GPU.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>

#include "../../include/sihl_od.h"

namespace {

constexpr int kK = 256;                       // in_features of every tower layer (the reference's num_channels default)
constexpr int kBlockM = 128;                  // rows per tile = TMEM lanes
constexpr int kChunkK = 64;                   // bf16 per 128-byte swizzle row
constexpr int kChunks = kK / kChunkK;         // 4
constexpr int kUmmaK = 16;                    // K per tcgen05.mma (bf16)
constexpr int kXStageBytes = kBlockM * 128;   // one ring slot: 128 rows x 128 B
constexpr int kThreads = 192;                 // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2-5: epilogue

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
        "selp.u32 %0, 1, 0, P1;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: ~2 s of polling, then trap (a launch failure the host sees) — never a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int what) {
    if (mbar_try(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) {
            printf("sihl_od_mlp: barrier wait timed out (block %d thread %d wait-site %d parity %u)\n", blockIdx.x, threadIdx.x, what, parity);
            __trap();
        }
    }
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n"
        ".reg .b32 rx;\n"
        ".reg .pred px;\n"
        "elect.sync rx|px, %1;\n"
        "@px mov.s32 %0, 1;\n"
        "}\n"
        : "+r"(pred)
        : "r"(0xFFFFFFFFu));
    return pred != 0;
}

__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, both operands K-major bf16, fp32 accumulate; issued by ONE thread.
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Shared-memory matrix descriptor: K-major tile of 128-byte rows, 128-byte swizzle, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t sw128_kmajor_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);     // start address, 16-byte units        bits [0,14)
    d |= static_cast<uint64_t>(1) << 16;                     // leading byte offset (unused here)   bits [16,30)
    d |= static_cast<uint64_t>(1024 >> 4) << 32;             // stride byte offset: 8 rows x 128 B  bits [32,46)
    d |= static_cast<uint64_t>(1) << 46;                     // descriptor version (Blackwell)      bits [46,48)
    d |= static_cast<uint64_t>(2) << 61;                     // SWIZZLE_128B                        bits [61,64)
    return d;
}
// Instruction descriptor of kind::f16: fp32 accumulator, bf16 A and B, both K-major, shape 128 x N.
template <int N>
__host__ __device__ constexpr uint32_t idesc_bf16_f32() {
    return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(kBlockM >> 4) << 24);
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        "tcgen05.wait::ld.sync.aligned;\n"          // same asm block: the registers are not read before the load has landed
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        "tcgen05.wait::ld.sync.aligned;\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}

__host__ __device__ constexpr int tmem_stage_cols(int n) { return n <= 32 ? 32 : n <= 64 ? 64 : n <= 128 ? 128 : 256; }

template <int N>
struct MlpSmem {
    static constexpr int kStages = (N == 256) ? 5 : 8;
    static constexpr int kWBytes = kChunks * N * 128;
    static constexpr int kRingBytes = kStages * kXStageBytes;
    static constexpr int kParamBytes = 3 * N * 4;
    static constexpr int kBarBytes = (2 * kStages + 1 + 4) * 8 + 16;
    static constexpr int kTotal = 1024 /* alignment slack */ + kWBytes + kRingBytes + kParamBytes + kBarBytes;
};

struct MlpParams {
    const float* bias;
    const float* gamma;      // hidden layers only
    const float* beta;       // hidden layers only
    void* out;               // hidden: bf16 [M,N]; output layer: fp32 [M,out_cols]
    long long M;
    int n_tiles;
    int out_cols;
    float eps;
};

// HIDDEN = true:  bias + LayerNorm + SiLU -> bf16 [M,N]   (N == 256)
// HIDDEN = false: bias                    -> fp32 [M,out_cols], out_cols <= N
template <int N, bool HIDDEN>
__global__ void __launch_bounds__(kThreads, 1)
k_mlp_layer(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const MlpParams p) {
    using L = MlpSmem<N>;
    constexpr int S = L::kStages;
    constexpr int kStageCols = tmem_stage_cols(N);
    constexpr int kTmemCols = 2 * kStageCols;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t* sW = base;                                   // 4 chunks x [N rows x 128 B], swizzled by the TMA unit
    uint8_t* sX = sW + L::kWBytes;                        // ring of [128 rows x 128 B] slots
    float* sBias = reinterpret_cast<float*>(sX + L::kRingBytes);
    float* sGamma = sBias + N;
    float* sBeta = sGamma + N;
    uint64_t* full = reinterpret_cast<uint64_t*>(sBeta + N);
    uint64_t* empty = full + S;
    uint64_t* w_full = empty + S;
    uint64_t* t_full = w_full + 1;                        // [2] accumulator stage ready for the epilogue
    uint64_t* t_empty = t_full + 2;                       // [2] accumulator stage drained
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    for (int i = threadIdx.x; i < N; i += kThreads) {
        sBias[i] = p.bias[i];
        sGamma[i] = HIDDEN ? p.gamma[i] : 1.f;
        sBeta[i] = HIDDEN ? p.beta[i] : 0.f;
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(w_full, 1);
        for (int s = 0; s < 2; ++s) { mbar_init(&t_full[s], 1); mbar_init(&t_empty[s], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {                                      // TMEM: one warp allocates and later frees
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (elect_one()) {
            mbar_expect_tx(w_full, L::kWBytes);
            for (int c = 0; c < kChunks; ++c) tma_load_2d(&map_w, w_full, sW + c * (N * 128), c * kChunkK, 0);
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
                for (int c = 0; c < kChunks; ++c, ++it) {
                    const uint32_t s = it % S, ph = (it / S) & 1;
                    mbar_wait(&empty[s], ph ^ 1, 0);
                    mbar_expect_tx(&full[s], kXStageBytes);
                    tma_load_2d(&map_x, &full[s], sX + s * kXStageBytes, c * kChunkK, tile * kBlockM);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (elect_one()) {
            constexpr uint32_t idesc = idesc_bf16_f32<N>();
            mbar_wait(w_full, 0, 1);
            uint32_t it = 0, t = 0;
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++t) {
                const uint32_t as = t & 1, aph = (t >> 1) & 1;
                mbar_wait(&t_empty[as], aph ^ 1, 2);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * kStageCols;
                for (int c = 0; c < kChunks; ++c, ++it) {
                    const uint32_t s = it % S, ph = (it / S) & 1;
                    mbar_wait(&full[s], ph, 3);
                    tc_fence_after();
                    const uint32_t a_base = smem_u32(sX + s * kXStageBytes);
                    const uint32_t b_base = smem_u32(sW + c * (N * 128));
#pragma unroll
                    for (int k = 0; k < kChunkK / kUmmaK; ++k) {
                        tc_mma_bf16(d_tmem, sw128_kmajor_desc(a_base + k * (kUmmaK * 2)), sw128_kmajor_desc(b_base + k * (kUmmaK * 2)), idesc,
                                    static_cast<uint32_t>((c | k) != 0));
                    }
                    tc_commit(&empty[s]);                 // ring slot free once these MMAs have read it
                }
                tc_commit(&t_full[as]);                   // accumulator complete
            }
        }
    } else {
        // ===== epilogue: warp w may touch TMEM lanes 32*(w%4) .. +31; thread = row =====
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        uint32_t t = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++t) {
            const uint32_t as = t & 1, aph = (t >> 1) & 1;
            mbar_wait(&t_full[as], aph, 4);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * kStageCols;
            const long long grow = static_cast<long long>(tile) * kBlockM + row;
            if constexpr (HIDDEN) {
                uint32_t v[32];
                float sum = 0.f;
#pragma unroll 1
                for (int c0 = 0; c0 < N; c0 += 32) {
                    tmem_ld32(taddr + c0, v);
#pragma unroll
                    for (int j = 0; j < 32; ++j) sum += __uint_as_float(v[j]) + sBias[c0 + j];
                }
                const float mean = sum * (1.f / N);
                float ssq = 0.f;
#pragma unroll 1
                for (int c0 = 0; c0 < N; c0 += 32) {
                    tmem_ld32(taddr + c0, v);
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float d = (__uint_as_float(v[j]) + sBias[c0 + j]) - mean;
                        ssq = __fmaf_rn(d, d, ssq);
                    }
                }
                const float rstd = 1.f / sqrtf(ssq * (1.f / N) + p.eps);
                __nv_bfloat16* orow = reinterpret_cast<__nv_bfloat16*>(p.out) + grow * N;
#pragma unroll 1
                for (int c0 = 0; c0 < N; c0 += 32) {
                    tmem_ld32(taddr + c0, v);
                    uint32_t packed[16];
#pragma unroll
                    for (int j = 0; j < 32; j += 2) {
                        float y[2];
#pragma unroll
                        for (int u = 0; u < 2; ++u) {
                            const float d = ((__uint_as_float(v[j + u]) + sBias[c0 + j + u]) - mean) * rstd;
                            const float z = __fmaf_rn(d, sGamma[c0 + j + u], sBeta[c0 + j + u]);
                            // SiLU(z) = z * sigmoid(z) = h + h * tanh(h), h = z / 2: one MUFU op per element
                            const float h = 0.5f * z;
                            float th;
                            asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(h));
                            y[u] = __fmaf_rn(h, th, h);
                        }
                        const __nv_bfloat162 b2 = __floats2bfloat162_rn(y[0], y[1]);
                        packed[j >> 1] = *reinterpret_cast<const uint32_t*>(&b2);
                    }
                    if (grow < p.M) {
                        uint4* dst = reinterpret_cast<uint4*>(orow + c0);
#pragma unroll
                        for (int q = 0; q < 4; ++q) dst[q] = make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
                    }
                }
            } else {
                uint32_t v[16];
                float* orow = reinterpret_cast<float*>(p.out) + grow * p.out_cols;
#pragma unroll 1
                for (int c0 = 0; c0 < N; c0 += 16) {
                    tmem_ld16(taddr + c0, v);
                    if (grow < p.M) {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (c0 + j < p.out_cols) orow[c0 + j] = __uint_as_float(v[j]) + sBias[c0 + j];
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&t_empty[as]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// ---- host side ---------------------------------------------------------------------------------------------------------
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = [] {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult status;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &status) != cudaSuccess || status != cudaDriverEntryPointSuccess)
            ptr = nullptr;
        return reinterpret_cast<EncodeTiledFn>(ptr);
    }();
    return fn;
}

// [rows, 256] bf16 row-major matrix, read in boxes of box_rows x 64 columns (128 B) with the 128-byte swizzle.
bool make_map(CUtensorMap* map, const void* ptr, unsigned long long rows, unsigned box_rows) {
    EncodeTiledFn fn = encode_tiled();
    if (!fn) return false;
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(kK), static_cast<cuuint64_t>(rows)};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(kK) * 2};
    const cuuint32_t box[2] = {static_cast<cuuint32_t>(kChunkK), box_rows};
    const cuuint32_t elem[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, elem, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int sm_count() {
    static int n = [] {
        int dev = 0, v = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) v = 0;
        return v;
    }();
    return n;
}

template <int N, bool HIDDEN>
int launch_layer(const void* x, long long M, const void* w, const MlpParams& p_in, cudaStream_t stream) {
    using L = MlpSmem<N>;
    // per launch: the attribute belongs to the current device's context, and a process may drive several
    if (cudaFuncSetAttribute(k_mlp_layer<N, HIDDEN>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal) != cudaSuccess) return SIHL_OD_ECUDA;
    CUtensorMap map_x, map_w;
    if (!make_map(&map_x, x, static_cast<unsigned long long>(M), kBlockM) || !make_map(&map_w, w, N, N)) return SIHL_OD_ECUDA;
    MlpParams p = p_in;
    p.M = M;
    p.n_tiles = static_cast<int>((M + kBlockM - 1) / kBlockM);
    const int sms = sm_count();
    if (sms <= 0) return SIHL_OD_ECUDA;
    const int grid = p.n_tiles < sms ? p.n_tiles : sms;
    k_mlp_layer<N, HIDDEN><<<grid, kThreads, L::kTotal, stream>>>(map_x, map_w, p);
    return cudaGetLastError() == cudaSuccess ? SIHL_OD_OK : SIHL_OD_ECUDA;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace

extern "C" {

SIHL_OD_API int sihl_od_mlp_hidden(const void* x_bf16, int64_t M, int channels, const void* w_bf16, const float* bias, const float* gamma,
                                   const float* beta, float eps, void* y_bf16, void* stream) {
    if (channels != kK || M < 0 || M > 0x7FFFFF00LL) return SIHL_OD_EINVAL;
    if (M == 0) return SIHL_OD_OK;
    if (!x_bf16 || !w_bf16 || !bias || !gamma || !beta || !y_bf16 || !aligned16(x_bf16) || !aligned16(w_bf16) || !aligned16(y_bf16))
        return SIHL_OD_EINVAL;
    MlpParams p{};
    p.bias = bias; p.gamma = gamma; p.beta = beta; p.out = y_bf16; p.eps = eps; p.out_cols = kK;
    return launch_layer<256, true>(x_bf16, M, w_bf16, p, static_cast<cudaStream_t>(stream));
}

SIHL_OD_API int sihl_od_mlp_out(const void* x_bf16, int64_t M, int channels, const void* w_bf16, const float* bias, int n_pad, int out_cols,
                                float* y, void* stream) {
    if (channels != kK || M < 0 || M > 0x7FFFFF00LL || out_cols < 1 || out_cols > n_pad) return SIHL_OD_EINVAL;
    if (n_pad != 16 && n_pad != 32 && n_pad != 64 && n_pad != 96 && n_pad != 128 && n_pad != 256) return SIHL_OD_EINVAL;
    if (M == 0) return SIHL_OD_OK;
    if (!x_bf16 || !w_bf16 || !bias || !y || !aligned16(x_bf16) || !aligned16(w_bf16)) return SIHL_OD_EINVAL;
    MlpParams p{};
    p.bias = bias; p.out = y; p.out_cols = out_cols;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (n_pad) {
        case 16: return launch_layer<16, false>(x_bf16, M, w_bf16, p, st);
        case 32: return launch_layer<32, false>(x_bf16, M, w_bf16, p, st);
        case 64: return launch_layer<64, false>(x_bf16, M, w_bf16, p, st);
        case 96: return launch_layer<96, false>(x_bf16, M, w_bf16, p, st);
        case 128: return launch_layer<128, false>(x_bf16, M, w_bf16, p, st);
        default: return launch_layer<256, false>(x_bf16, M, w_bf16, p, st);
    }
}

}  // extern "C"

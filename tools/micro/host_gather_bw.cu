// Developer microbenchmark (not part of the product): how fast can a kernel GATHER rows from pinned host memory over PCIe?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -cudart shared -o /tmp/host_gather_bw host_gather_bw.cu && /tmp/host_gather_bw
// The end-to-end path of bench.py leaves the class map in pinned host memory and reads the positives' rows in place
// (38 470 rows of 320 B out of 545 600 at cfg1).  Variants: (1) LSU loads, 8 lanes x 16 B per 128-byte segment (what
// k_pos_loss_tiles does), (2) LSU loads, 20 lanes x 16 B = the whole row per warp instruction, (3) one cp.async.bulk of 320 B
// per row into shared memory (TMA unit), N rows in flight per warp, (4) contiguous streaming of the same number of bytes.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <algorithm>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
constexpr int kRowBytes = 320;

__global__ void __launch_bounds__(256) k_lsu8(const char *base, const int *rows, int n, float *out)
{
    float acc = 0.f;
    const int group = (blockIdx.x * blockDim.x + threadIdx.x) >> 3, gl = threadIdx.x & 7, groups = (gridDim.x * blockDim.x) >> 3;
    for (int i = group; i < n; i += groups) {
        const float4 *p = reinterpret_cast<const float4 *>(base + (size_t)rows[i] * kRowBytes);
        float4 a = p[gl], b = p[gl + 8], c = make_float4(0, 0, 0, 0);
        if (gl < 4) c = p[gl + 16];
        acc += a.x + b.y + c.z;
    }
    if (acc == 123.456f) out[0] = acc;
}

__global__ void __launch_bounds__(256) k_lsu20(const char *base, const int *rows, int n, float *out)
{
    float acc = 0.f;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31, warps = (gridDim.x * blockDim.x) >> 5;
    for (int i = warp * 4; i < n; i += warps * 4) {
        float4 v[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            v[r] = make_float4(0, 0, 0, 0);
            if (i + r < n && lane < 20) v[r] = reinterpret_cast<const float4 *>(base + (size_t)rows[i + r] * kRowBytes)[lane];
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) acc += v[r].x + v[r].w;
    }
    if (acc == 123.456f) out[0] = acc;
}

__device__ __forceinline__ uint32_t smem_u32(const void *ptr) { return (uint32_t)__cvta_generic_to_shared(ptr); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t *bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(b) : "memory"); }
__device__ __forceinline__ void bulk(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}

// every warp: batches of ROWS rows, one bulk copy per row issued by lanes 0..ROWS-1, one mbarrier per warp
template <int ROWS>
__global__ void __launch_bounds__(128) k_bulk(const char *base, const int *rows, int n, float *out)
{
    extern __shared__ __align__(128) unsigned char s[];
    const int wi = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *buf = s + (size_t)wi * ROWS * kRowBytes;
    uint64_t *bar = reinterpret_cast<uint64_t *>(s + (size_t)4 * ROWS * kRowBytes) + wi;
    if (lane == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncwarp();
    const int warp = blockIdx.x * 4 + wi, warps = gridDim.x * 4;
    float acc = 0.f;
    uint32_t ph = 0;
    for (int i = warp * ROWS; i < n; i += warps * ROWS) {
        const int cnt = min(ROWS, n - i);
        if (lane == 0) mbar_expect(bar, (uint32_t)cnt * kRowBytes);
        __syncwarp();
        if (lane < cnt) bulk(buf + lane * kRowBytes, base + (size_t)rows[i + lane] * kRowBytes, kRowBytes, bar);
        long long t0 = clock64();
        while (!mbar_try(bar, ph)) { if (clock64() - t0 > 4000000000LL) { __trap(); } }
        ph ^= 1;
        for (int j = lane; j < cnt * (kRowBytes / 16); j += 32) { const float4 v = reinterpret_cast<const float4 *>(buf)[j]; acc += v.x + v.w; }
        __syncwarp();
    }
    if (acc == 123.456f) out[0] = acc;
}

__global__ void __launch_bounds__(256) k_stream(const float4 *p, size_t n4, float *out)
{
    float acc = 0.f;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) { const float4 v = p[i]; acc += v.x + v.w; }
    if (acc == 123.456f) out[0] = acc;
}

int main()
{
    const int total_rows = 545600, n = 38470, reps = 10;
    char *h = nullptr;
    CK(cudaHostAlloc(&h, (size_t)total_rows * kRowBytes, cudaHostAllocDefault));
    for (size_t i = 0; i < (size_t)total_rows * kRowBytes; i += 4096) h[i] = (char)i;
    std::vector<int> idx(total_rows);
    for (int i = 0; i < total_rows; ++i) idx[i] = i;
    srand(1);
    for (int i = 0; i < n; ++i) std::swap(idx[i], idx[i + rand() % (total_rows - i)]);
    std::sort(idx.begin(), idx.begin() + n);
    int *d_rows; float *d_out;
    CK(cudaMalloc(&d_rows, n * sizeof(int))); CK(cudaMalloc(&d_out, 4));
    CK(cudaMemcpy(d_rows, idx.data(), n * sizeof(int), cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const double mb = (double)n * kRowBytes / 1e6;
    auto report = [&](const char *name, float ms) { printf("%-44s %8.1f us  %6.1f GB/s\n", name, ms * 1e3 / reps, mb / (ms / reps)); };
    float ms;
    for (int grid : {148, 296, 592, 1184}) {
        k_lsu8<<<grid, 256>>>(h, d_rows, n, d_out); CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0)); for (int r = 0; r < reps; ++r) k_lsu8<<<grid, 256>>>(h, d_rows, n, d_out); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1)); char nm[64]; snprintf(nm, 64, "lsu 8 lanes x 16 B, grid %d", grid); report(nm, ms);
    }
    for (int grid : {148, 296, 592, 1184}) {
        k_lsu20<<<grid, 256>>>(h, d_rows, n, d_out); CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0)); for (int r = 0; r < reps; ++r) k_lsu20<<<grid, 256>>>(h, d_rows, n, d_out); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1)); char nm[64]; snprintf(nm, 64, "lsu 20 lanes x 16 B x 4 rows, grid %d", grid); report(nm, ms);
    }
#define BULK(ROWS)                                                                                                            \
    for (int grid : {148, 296, 592}) {                                                                                        \
        const size_t smem = (size_t)4 * ROWS * kRowBytes + 64;                                                                \
        CK(cudaFuncSetAttribute(k_bulk<ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                         \
        k_bulk<ROWS><<<grid, 128, smem>>>(h, d_rows, n, d_out); CK(cudaDeviceSynchronize());                                   \
        CK(cudaEventRecord(e0)); for (int r = 0; r < reps; ++r) k_bulk<ROWS><<<grid, 128, smem>>>(h, d_rows, n, d_out);         \
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));                           \
        char nm[64]; snprintf(nm, 64, "cp.async.bulk 320 B, %d rows/warp, grid %d", ROWS, grid); report(nm, ms);               \
    }
    BULK(8) BULK(16) BULK(32)
    {
        const size_t n4 = (size_t)n * kRowBytes / 16;
        k_stream<<<592, 256>>>(reinterpret_cast<const float4 *>(h), n4, d_out); CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0)); for (int r = 0; r < reps; ++r) k_stream<<<592, 256>>>(reinterpret_cast<const float4 *>(h) + (size_t)r * n4, n4, d_out);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1)); report("contiguous LDG.128 stream (same bytes)", ms);
        char *d = nullptr; CK(cudaMalloc(&d, (size_t)n * kRowBytes * reps));
        CK(cudaEventRecord(e0)); CK(cudaMemcpyAsync(d, h, (size_t)n * kRowBytes * reps, cudaMemcpyHostToDevice)); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1)); report("copy engine, contiguous (same bytes x reps)", ms);
    }
    return 0;
}

// Developer microbenchmark (not part of the product): what read bandwidth can a streaming kernel reach on this GPU?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -cudart shared -o /tmp/read_bw read_bw.cu && /tmp/read_bw   (binary outside the tree)
// (1) plain LDG.128 grid-stride reduction, (2) the cp.async.bulk ring of k_dense_decode_tma with an empty consumer,
// (3) the same ring with one LDS.128 pass + max tree over the stage (the decode kernel's consumer without the argmax).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void __launch_bounds__(256) k_ldg(const float4 *p, size_t n4, float *out, int unroll_dummy)
{
    float acc = 0.f;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 7 * stride < n4; i += 8 * stride) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldcs(p + i + u * stride);
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
    }
    for (; i < n4; i += stride) { float4 v = __ldcs(p + i); acc += v.x + v.y + v.z + v.w; }
    if (acc == 123.456f) out[0] = acc;
}

__device__ __forceinline__ uint32_t smem_u32(const void *ptr) { return (uint32_t)__cvta_generic_to_shared(ptr); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t *bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(b) : "memory"); }
__device__ __forceinline__ void bulk(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile("{\n.reg .pred P1;\nW: mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra D;\nbra W;\nD:\n}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

template <int WORK>
__global__ void __launch_bounds__(256) k_ring(const char *p, size_t bytes, int stage_bytes, int stages, float *out)
{
    extern __shared__ __align__(128) unsigned char s[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(s + (size_t)stages * stage_bytes);
    const int tid = threadIdx.x;
    const int n_chunks = (int)(bytes / stage_bytes);
    if (tid == 0) { for (int i = 0; i < stages; ++i) mbar_init(bars + i, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    if (tid == 0) for (int i = 0; i < stages; ++i) { const int c = blockIdx.x + i * gridDim.x; if (c < n_chunks) { mbar_expect(bars + i, stage_bytes); bulk(s + (size_t)i * stage_bytes, p + (size_t)c * stage_bytes, stage_bytes, bars + i); } }
    float acc = 0.f;
    int st = 0; uint32_t ph = 0;
    for (int c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        mbar_wait(bars + st, ph);
        if (WORK) {
            const float4 *q = reinterpret_cast<const float4 *>(s + (size_t)st * stage_bytes);
            for (int i = tid; i < stage_bytes / 16; i += 256) { float4 v = q[i]; acc = fmaxf(acc, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w))); }
        }
        __syncthreads();
        if (tid == 0) { const int nx = c + stages * gridDim.x; if (nx < n_chunks) { mbar_expect(bars + st, stage_bytes); bulk(s + (size_t)st * stage_bytes, p + (size_t)nx * stage_bytes, stage_bytes, bars + st); } }
        if (++st == stages) { st = 0; ph ^= 1; }
    }
    if (acc == 123.456f) out[0] = acc;
}

int main()
{
    const size_t bytes = (size_t)186 << 20;               // one cfg1 step's class + box + loc maps
    const int n_buf = 3;                                    // rotate buffers: > 126 MB L2
    char *buf[n_buf]; float *out;
    for (int i = 0; i < n_buf; ++i) { CK(cudaMalloc(&buf[i], bytes)); CK(cudaMemset(buf[i], 1, bytes)); }
    CK(cudaMalloc(&out, 4));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto report = [&](const char *name, float ms, int iters) { printf("%-44s %8.2f us  %7.1f GB/s\n", name, ms / iters * 1e3, bytes / (ms / iters * 1e-3) / 1e9); };
    const int iters = 30;
    for (int ctas : {148 * 2, 148 * 4, 148 * 8}) {
        for (int w = 0; w < 3; ++w) k_ldg<<<ctas, 256>>>((const float4 *)buf[w % n_buf], bytes / 16, out, 0);
        cudaEventRecord(e0);
        for (int i = 0; i < iters; ++i) k_ldg<<<ctas, 256>>>((const float4 *)buf[i % n_buf], bytes / 16, out, 0);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        char name[64]; snprintf(name, 64, "LDG.128 x8 in flight, %d CTAs", ctas); report(name, ms, iters);
    }
    for (int work = 0; work < 2; ++work)
        for (int stage_kb : {21, 42})
            for (int stages : {2, 3, 4})
                for (int per_sm : {1, 2, 3}) {
                    const int stage_bytes = stage_kb * 1024;
                    const size_t smem = (size_t)stages * stage_bytes + 64;
                    if (smem * per_sm > 220 * 1024) continue;
                    auto kern = work ? k_ring<1> : k_ring<0>;
                    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    for (int w = 0; w < 3; ++w) kern<<<148 * per_sm, 256, smem>>>(buf[w % n_buf], bytes, stage_bytes, stages, out);
                    cudaEventRecord(e0);
                    for (int i = 0; i < iters; ++i) kern<<<148 * per_sm, 256, smem>>>(buf[i % n_buf], bytes, stage_bytes, stages, out);
                    cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
                    float ms; cudaEventElapsedTime(&ms, e0, e1);
                    char name[96]; snprintf(name, 96, "bulk ring %s %d KB x %d stages, %d CTA/SM", work ? "+LDS max" : "empty   ", stage_kb, stages, per_sm);
                    report(name, ms, iters);
                }
    return 0;
}

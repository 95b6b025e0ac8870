// od_common.cuh — shared device helpers for libsihl_b200 (sm_100a only).
//
// Arithmetic contract: every kernel that produces values compared bit-for-bit
// with torch-CUDA eager (anchors, CIoU, decoded boxes, sigmoid) is compiled with
// -fmad=false -prec-div=true -prec-sqrt=true and never --use_fast_math, so each
// source operator is one IEEE fp32 operation, in the operator order of the
// torch / torchvision eager code.  Fused multiply-adds appear
// only where torch's own kernels have them, written explicitly (__fmaf_rn).
#pragma once

#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "../../include/sihl_od.h"
#include "od_math.h"

namespace sihl {

void set_error(const char *fmt, ...);
int cuda_status(cudaError_t e, const char *what);

#define SIHL_CHECK_ARG(cond, ...)                          \
    do {                                                   \
        if (!(cond)) {                                     \
            ::sihl::set_error(__VA_ARGS__);                \
            return SIHL_OD_EINVAL;                         \
        }                                                  \
    } while (0)

#define SIHL_CHECK_LAUNCH(what)                                              \
    do {                                                                     \
        int _rc = ::sihl::cuda_status(cudaGetLastError(), what);             \
        if (_rc) return _rc;                                                 \
    } while (0)

constexpr int kNumSMs = 148;          // B200
constexpr unsigned kFullMask = 0xffffffffu;
constexpr int kTile = 512;            // anchors per CTA of k_assign_resolve == capacity of one tile's positive list

struct LevelTable {                   // passed by value as a kernel parameter
    int n;
    int h[SIHL_OD_MAX_LEVELS];
    int w[SIHL_OD_MAX_LEVELS];
    int base[SIHL_OD_MAX_LEVELS + 1]; // first anchor index of each level
};

__device__ __forceinline__ float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }

__device__ __forceinline__ Box4 to_box(float4 b) { return Box4{b.x, b.y, b.z, b.w}; }

template <typename T>
__device__ __forceinline__ T warp_sum(T v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
    return v;
}

__device__ __forceinline__ float warp_max(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFullMask, v, o));
    return v;
}

// binary_cross_entropy_with_logits term (od_math.h bce_logits) for the fused resolve pass, where it is 13 % of the
// kernel's instructions: log1p(e^-|x|) with y = e^-|x| from ex2.approx and, for y < 0.223 (|x| > 1.5: all but a
// handful of the locations of a detection head whose location logits sit around -5), the alternating series
// y - y^2/2 + ... - y^8/8 (truncation <= y^8/9 = 7e-7 relative) in 8 FMAs; the exact libdevice path otherwise.
// The location loss is held to 1e-5 relative (north_star), not to bit equality; sihl_od_dense_loss keeps the
// libdevice form.
__device__ __forceinline__ float bce_logits_fast(float x, float t)
{
    const float ax = fabsf(x);
    const float y = __expf(-ax);
    float l;
    if (y < 0.223f) {
        float q = -0.125f;
        q = __fmaf_rn(q, y, 0.14285714285714285f);
        q = __fmaf_rn(q, y, -0.16666666666666666f);
        q = __fmaf_rn(q, y, 0.2f);
        q = __fmaf_rn(q, y, -0.25f);
        q = __fmaf_rn(q, y, 0.33333333333333333f);
        q = __fmaf_rn(q, y, -0.5f);
        q = __fmaf_rn(q, y, 1.f);
        l = q * y;
    } else {
        l = log1pf(expf(-ax));
    }
    return (1.f - t) * x + (fmaxf(-x, 0.f) + l);
}

// Block-wide sum of NV doubles per thread -> atomicAdd into dst[slot[i]] by one thread.
// red: shared double[NV * 32].
template <int NV>
__device__ __forceinline__ void block_accumulate(double (&v)[NV], double *red, double *dst, const int (&slot)[NV])
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) red[i * 32 + warp] = v[i];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            double x = lane < nwarps ? red[i * 32 + lane] : 0.0;
            x = warp_sum(x);
            if (lane == 0 && x != 0.0) atomicAdd(dst + slot[i], x);
        }
    }
    __syncthreads();
}


// Block-wide sum of NV floats per thread (each the sum of a handful of elements): fp32 inside a warp,
// fp64 across warps and into dst[slot[i]] — one shared-memory hop, one atomicAdd per CTA and term.
// red: shared double[NV * 32].
template <int NV>
__device__ __forceinline__ void block_accumulate_f(float (&v)[NV], double *red, double *dst, const int (&slot)[NV])
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) red[i * 32 + warp] = (double)v[i];
    }
    __syncthreads();
    // lane i of warp 0 finishes term i: a handful of fp64 adds, then one atomicAdd per CTA and term
    if (warp == 0 && lane < NV) {
        double x = 0.0;
        for (int w = 0; w < nwarps; ++w) x += red[lane * 32 + w];
        int sl = 0;
#pragma unroll
        for (int i = 0; i < NV; ++i)
            if (lane == i) sl = slot[i];
        if (x != 0.0) atomicAdd(dst + sl, x);
    }
}

}  // namespace sihl

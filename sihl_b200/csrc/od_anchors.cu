// od_anchors.cu — anchor grid.
//
// Replaces ObjectDetection.get_offsets_and_scales (ref: src/sihl/heads/object_detection.py:83-97)
// and anchors = (offsets + scales) * full_size (ref :134-140): ~40 tiny ATen launches per call
// in the reference, one launch here.  One thread per anchor, float4 stores (HBM-write bound,
// 48 B per anchor; the tables are constants of the feature-map shapes and are cached by the
// host side, so this kernel is off the steady-state path).
#include "od_common.cuh"

namespace sihl {

struct AnchorParams {
    LevelTable lv;
    float start_x[SIHL_OD_MAX_LEVELS], end_x[SIHL_OD_MAX_LEVELS], step_x[SIHL_OD_MAX_LEVELS];
    float start_y[SIHL_OD_MAX_LEVELS], end_y[SIHL_OD_MAX_LEVELS], step_y[SIHL_OD_MAX_LEVELS];
    float half_x[SIHL_OD_MAX_LEVELS], half_y[SIHL_OD_MAX_LEVELS];
    float img_w, img_h;
};

// torch.linspace as ATen's CUDA kernel evaluates it (RangeFactories.cu): first half counts up
// from start, second half counts down from end, each a single fused multiply-add.
__device__ __forceinline__ float linspace_at(float start, float end, float step, int steps, int i)
{
    if (steps == 1) return start;
    return (i < steps / 2) ? __fmaf_rn(step, (float)i, start) : __fmaf_rn(-step, (float)(steps - i - 1), end);
}

__global__ void __launch_bounds__(256) k_anchors(AnchorParams p, float4 *__restrict__ offsets,
                                                 float4 *__restrict__ scales, float4 *__restrict__ anchors)
{
    const int total = p.lv.base[p.lv.n];
    for (int a = blockIdx.x * blockDim.x + threadIdx.x; a < total; a += gridDim.x * blockDim.x) {
        int l = 0;
#pragma unroll
        for (int k = 1; k < SIHL_OD_MAX_LEVELS; ++k) l += (k < p.lv.n && a >= p.lv.base[k]);
        const int r = a - p.lv.base[l], i = r / p.lv.w[l], j = r - i * p.lv.w[l];
        const float x = linspace_at(p.start_x[l], p.end_x[l], p.step_x[l], p.lv.w[l], j);
        const float y = linspace_at(p.start_y[l], p.end_y[l], p.step_y[l], p.lv.h[l], i);
        const float4 off = make_float4(x, y, x, y);
        const float4 sc = make_float4(-p.half_x[l], -p.half_y[l], p.half_x[l], p.half_y[l]);
        if (offsets) offsets[a] = off;
        if (scales) scales[a] = sc;
        if (anchors)
            anchors[a] = make_float4((off.x + sc.x) * p.img_w, (off.y + sc.y) * p.img_h,
                                     (off.z + sc.z) * p.img_w, (off.w + sc.w) * p.img_h);
    }
}

int fill_level_table(const int32_t *level_hw_host, int n_levels, LevelTable *lv)
{
    SIHL_CHECK_ARG(level_hw_host != nullptr, "level table is NULL");
    SIHL_CHECK_ARG(n_levels >= 1 && n_levels <= SIHL_OD_MAX_LEVELS, "n_levels=%d outside 1..%d", n_levels,
                   SIHL_OD_MAX_LEVELS);
    lv->n = n_levels;
    int64_t base = 0;
    for (int l = 0; l < SIHL_OD_MAX_LEVELS; ++l) {
        int h = l < n_levels ? level_hw_host[2 * l] : 1, w = l < n_levels ? level_hw_host[2 * l + 1] : 1;
        SIHL_CHECK_ARG(h > 0 && w > 0, "level %d has size %dx%d", l, h, w);
        lv->h[l] = h; lv->w[l] = w; lv->base[l] = (int)base;
        if (l < n_levels) base += (int64_t)h * w;
        SIHL_CHECK_ARG(base < (1ll << 30), "too many anchors");
    }
    for (int l = n_levels; l <= SIHL_OD_MAX_LEVELS; ++l) lv->base[l] = (int)base;
    lv->base[n_levels] = (int)base;
    return SIHL_OD_OK;
}

}  // namespace sihl

extern "C" int sihl_od_anchors(const int32_t *level_hw_host, int n_levels, int img_w, int img_h, float *offsets,
                               float *scales, float *anchors, void *stream)
{
    using namespace sihl;
    AnchorParams p;
    int rc = fill_level_table(level_hw_host, n_levels, &p.lv);
    if (rc) return rc;
    SIHL_CHECK_ARG(img_w > 0 && img_h > 0, "image size %dx%d", img_w, img_h);
    for (int l = 0; l < SIHL_OD_MAX_LEVELS; ++l) {
        const int h = p.lv.h[l], w = p.lv.w[l];
        // ref :88 — python doubles; they reach fp32 through linspace's scalar conversion and torch.tensor(...)
        const double y_min = 1.0 / h / 2.0, x_min = 1.0 / w / 2.0;
        p.start_x[l] = (float)x_min; p.end_x[l] = (float)(1.0 - x_min);
        p.start_y[l] = (float)y_min; p.end_y[l] = (float)(1.0 - y_min);
        p.step_x[l] = w > 1 ? (p.end_x[l] - p.start_x[l]) / (float)(w - 1) : 0.f;
        p.step_y[l] = h > 1 ? (p.end_y[l] - p.start_y[l]) / (float)(h - 1) : 0.f;
        p.half_x[l] = (float)x_min; p.half_y[l] = (float)y_min;
    }
    p.img_w = (float)img_w; p.img_h = (float)img_h;   // int64 tensor promoted to fp32, ref :134-136
    const int total = p.lv.base[n_levels];
    const int blocks = (total + 255) / 256;
    k_anchors<<<blocks < 1 ? 1 : blocks, 256, 0, (cudaStream_t)stream>>>(
        p, reinterpret_cast<float4 *>(offsets), reinterpret_cast<float4 *>(scales), reinterpret_cast<float4 *>(anchors));
    SIHL_CHECK_LAUNCH("k_anchors");
    return SIHL_OD_OK;
}

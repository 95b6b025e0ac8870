// od_map.cu — N3: the detection <-> ground-truth matching behind the validation mAP.
//
// The reference hands every validation batch to torchmetrics' MeanAveragePrecision (ref object_detection.py:219-237,
// :245; backend faster_coco_eval), whose compute() runs COCOeval on the CPU at epoch end: per image and category,
// detections in score order are greedily matched to the ground truth at the 10 IoU thresholds .50:.05:.95 and for
// the 4 area ranges (COCOeval.evaluateImg), then precision / recall curves are accumulated.  The matching is the
// part that scales with batch x detections x ground truth; it runs here, on the GPU, right after forward() —
// one CTA per image, one thread per (area range, IoU threshold) greedy chain, box IoU in fp64 as maskUtils.iou
// computes it on the xywh boxes torchmetrics builds (w = fl32(x2 - x1), h = fl32(y2 - y1)).
//
// Matching rule restated from pycocotools / faster_coco_eval COCOeval.evaluateImg (iscrowd == 0 everywhere):
//   gts of the category in input order, "ignored" ones (area outside the range) after the others;
//   for each detection d by (score desc, input order):  best = min(t, 1 - 1e-10), m = none
//       for each gt g in that order: skip g if already matched at this threshold;
//           stop if m is a regular gt and g is an ignored one; skip if iou(d, g) < best; best = iou(d, g), m = g
//       if m: d is matched to m (and inherits its ignore flag), m is taken
//       else: d is unmatched, and ignored iff its own area is outside the range.
// torchmetrics / faster_coco_eval are not installed in this environment: parity of this row is UNPINNED (the oracle,
// oracle/map_oracle.py, restates the same published algorithm).
#include "od_common.cuh"

namespace sihl {

constexpr int kMapMaxT = 16, kMapMaxA = 8, kMapMaxK = 1024;

struct MapParams {
    const float4 *det_boxes; const float *det_scores; const int64_t *det_classes; int K;
    const float4 *gt_boxes; const int64_t *gt_classes; const int32_t *gt_offsets;
    int n_thr, n_area;
    double thr[kMapMaxT];
    double area_lo[kMapMaxA], area_hi[kMapMaxA];
    int32_t *det_order;        // [B, K] rank -> detection
    int32_t *dt_match;         // [B, n_area, n_thr, K] by rank: global gt index or -1
    uint8_t *dt_ignore;        // [B, n_area, n_thr, K] by rank
    uint8_t *gt_ignore;        // [n_area, total_gt]
    int32_t *gt_taken;         // workspace [n_area, n_thr, total_gt]: 0 or 1 + rank of the detection that took the gt
    int total_gt;
};

__device__ __forceinline__ double box_area_d(float4 b)
{
    return (double)(b.z - b.x) * (double)(b.w - b.y);            // w, h rounded to fp32 first (torchmetrics box_convert)
}

// maskUtils.iou on xywh boxes (bbIou, iscrowd = 0), in double.
__device__ __forceinline__ double box_iou_d(float4 d, float4 g)
{
    const double dx = d.x, dy = d.y, dw = (double)(d.z - d.x), dh = (double)(d.w - d.y);
    const double gx = g.x, gy = g.y, gw = (double)(g.z - g.x), gh = (double)(g.w - g.y);
    const double w = fmin(dx + dw, gx + gw) - fmax(dx, gx);
    if (w <= 0.0) return 0.0;
    const double h = fmin(dy + dh, gy + gh) - fmax(dy, gy);
    if (h <= 0.0) return 0.0;
    const double i = w * h;
    const double u = dw * dh + gw * gh - i;
    return i / u;
}

__global__ void __launch_bounds__(128) k_map_match(const __grid_constant__ MapParams p)
{
    __shared__ int s_order[kMapMaxK];
    const int b = blockIdx.x, tid = threadIdx.x, K = p.K;
    const float *sc = p.det_scores + (int64_t)b * K;
    const int64_t *dc = p.det_classes + (int64_t)b * K;
    const float4 *db = p.det_boxes + (int64_t)b * K;
    const int g0 = __ldg(p.gt_offsets + b), g1 = __ldg(p.gt_offsets + b + 1);

    // detections by (score desc, input order): rank by counting (K <= 1024)
    for (int d = tid; d < K; d += blockDim.x) {
        const float s = __ldg(sc + d);
        int r = 0;
        for (int j = 0; j < K; ++j) {
            const float sj = __ldg(sc + j);
            r += (sj > s) || (sj == s && j < d);
        }
        s_order[r] = d;
        p.det_order[(int64_t)b * K + r] = d;
    }
    // ignore flags of this image's gts per area range; clear the "taken" markers
    for (int i = tid; i < (g1 - g0) * p.n_area; i += blockDim.x) {
        const int a = i / (g1 - g0), g = g0 + i % (g1 - g0);
        const double ar = box_area_d(__ldg(p.gt_boxes + g));
        p.gt_ignore[(int64_t)a * p.total_gt + g] = (ar < p.area_lo[a] || ar > p.area_hi[a]) ? 1 : 0;
    }
    for (int i = tid; i < (g1 - g0) * p.n_area * p.n_thr; i += blockDim.x)
        p.gt_taken[(int64_t)(i / (g1 - g0)) * p.total_gt + g0 + i % (g1 - g0)] = 0;
    __syncthreads();

    // one greedy chain per (area range, threshold)
    for (int chain = tid; chain < p.n_area * p.n_thr; chain += blockDim.x) {
        const int a = chain / p.n_thr, t = chain % p.n_thr;
        const uint8_t *gig = p.gt_ignore + (int64_t)a * p.total_gt;
        int32_t *taken = p.gt_taken + (int64_t)chain * p.total_gt;
        int32_t *dtm = p.dt_match + (((int64_t)b * p.n_area + a) * p.n_thr + t) * K;
        uint8_t *dti = p.dt_ignore + (((int64_t)b * p.n_area + a) * p.n_thr + t) * K;
        const double thr = fmin(p.thr[t], 1.0 - 1e-10);
        for (int r = 0; r < K; ++r) {
            const int d = s_order[r];
            const int64_t cls = __ldg(dc + d);
            const float4 box = __ldg(db + d);
            double best = thr;
            int m = -1;
            for (int pass = 0; pass < 2 && m < 0; ++pass) {        // regular gts first; ignored ones only while unmatched
                for (int g = g0; g < g1; ++g) {
                    if (__ldg(p.gt_classes + g) != cls || gig[g] != pass || taken[g] != 0) continue;
                    const double iou = box_iou_d(box, __ldg(p.gt_boxes + g));
                    if (iou < best) continue;
                    best = iou;
                    m = g;
                }
            }
            if (m >= 0) {
                dtm[r] = m;
                dti[r] = gig[m];
                taken[m] = r + 1;
            } else {
                const double ar = box_area_d(box);
                dtm[r] = -1;
                dti[r] = (ar < p.area_lo[a] || ar > p.area_hi[a]) ? 1 : 0;
            }
        }
    }
}

}  // namespace sihl

using namespace sihl;

extern "C" size_t sihl_od_map_workspace_bytes(int total_gt, int n_thresholds, int n_areas)
{
    if (total_gt < 0 || n_thresholds < 1 || n_areas < 1) return 0;
    return (size_t)(total_gt > 0 ? total_gt : 1) * (size_t)n_thresholds * (size_t)n_areas * sizeof(int32_t);
}

extern "C" int sihl_od_map_match(const float *det_boxes, const float *det_scores, const int64_t *det_classes, int batch, int k,
                                 const float *gt_boxes, const int64_t *gt_classes, const int32_t *gt_offsets, int total_gt,
                                 const double *iou_thresholds_host, int n_thresholds, const double *area_ranges_host,
                                 int n_areas, int32_t *det_order, int32_t *dt_match, uint8_t *dt_ignore, uint8_t *gt_ignore,
                                 void *workspace, void *stream)
{
    SIHL_CHECK_ARG(det_boxes && det_scores && det_classes && gt_offsets && iou_thresholds_host && area_ranges_host, "NULL input");
    SIHL_CHECK_ARG(det_order && dt_match && dt_ignore && workspace, "NULL output / workspace");
    SIHL_CHECK_ARG(batch >= 0 && k >= 1 && k <= kMapMaxK && total_gt >= 0, "bad sizes (k=%d, at most %d detections per image)", k,
                   kMapMaxK);
    SIHL_CHECK_ARG(n_thresholds >= 1 && n_thresholds <= kMapMaxT && n_areas >= 1 && n_areas <= kMapMaxA,
                   "at most %d IoU thresholds and %d area ranges", kMapMaxT, kMapMaxA);
    SIHL_CHECK_ARG(total_gt == 0 || (gt_boxes && gt_classes && gt_ignore), "NULL ground truth");
    if (batch == 0) return SIHL_OD_OK;
    MapParams p;
    p.det_boxes = reinterpret_cast<const float4 *>(det_boxes); p.det_scores = det_scores; p.det_classes = det_classes; p.K = k;
    p.gt_boxes = reinterpret_cast<const float4 *>(gt_boxes); p.gt_classes = gt_classes; p.gt_offsets = gt_offsets;
    p.n_thr = n_thresholds; p.n_area = n_areas;
    for (int i = 0; i < n_thresholds; ++i) p.thr[i] = iou_thresholds_host[i];
    for (int i = 0; i < n_areas; ++i) { p.area_lo[i] = area_ranges_host[2 * i]; p.area_hi[i] = area_ranges_host[2 * i + 1]; }
    p.det_order = det_order; p.dt_match = dt_match; p.dt_ignore = dt_ignore; p.gt_ignore = gt_ignore;
    p.gt_taken = reinterpret_cast<int32_t *>(workspace); p.total_gt = total_gt;
    k_map_match<<<batch, 128, 0, (cudaStream_t)stream>>>(p);
    SIHL_CHECK_LAUNCH("k_map_match");
    return SIHL_OD_OK;
}

/*
 * od_oracle.c — CPU restatement of sihl's ObjectDetection-head dense tail.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under sihl_b200/ may import, link or call
 * this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and there only as the checker / the CPU arm.
 *
 * Every function cites the reference lines it follows.  "ref:" paths are
 * relative to /root/reference, "tv:" paths are relative to the installed
 * torchvision 0.26 package (the arithmetic of complete_box_iou /
 * complete_box_iou_loss / nms lives there, not in the reference tree).
 *
 * Parity pinning: the reference's own tests hold no golden vectors for this
 * path.  The oracle is therefore pinned against outputs of the
 * reference itself, generated in the authoring container by
 * oracle/make_golden.py (imports /root/reference) and committed under
 * tests/golden/.  tests/test_oracle_golden.py checks every fixture.
 * NMS has no reference code at all (north-star extension): its oracle follows
 * torchvision's per-class path (tv:ops/boxes.py:102-120 + the CPU nms kernel)
 * and is pinned against torchvision outputs only — "parity unpinned" w.r.t.
 * the reference.
 *
 * Arithmetic rules: fp32, one IEEE operation per source operator, in the
 * operator order of the torch / torchvision eager code.  Build with
 * -ffp-contract=off (see oracle/Makefile); never -ffast-math.
 * Loss *sums* are accumulated in double (the reference uses fp32 tree sums;
 * the contract for losses is 1e-5 relative, not bit equality).
 */
#define _GNU_SOURCE
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_PI 3.14159265358979323846 /* torch.pi */
#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ */
/* a1/a2: anchor grid.  ref: src/sihl/heads/object_detection.py:83-97, */
/* :134-140.                                                           */
/* ------------------------------------------------------------------ */

/* torch.linspace(start, end, steps) in fp32 as ATen's CUDA kernel computes it
 * (aten/src/ATen/native/cuda/RangeFactories.cu, linspace_cuda_out): step is an
 * fp32 quotient, the first half counts up from start, the second half counts
 * down from end; nvcc contracts a+b*c to one fma.  The CPU ATen kernel uses a
 * vectorised arange whose last bits depend on the SIMD width, which is why the
 * golden fixtures store the reference's own tables and this function is pinned
 * to them only to 1e-6 relative (tests/test_oracle_golden.py). */
static void orc_linspace(double start_d, double end_d, int steps, float *out)
{
    float start = (float)start_d, end = (float)end_d;
    if (steps == 1) { out[0] = start; return; }
    float step = (end - start) / (float)(steps - 1);
    int halfway = steps / 2;
    for (int i = 0; i < steps; ++i) {
        if (i < halfway) out[i] = fmaf(step, (float)i, start);
        else             out[i] = fmaf(-step, (float)(steps - i - 1), end);
    }
}

/* level_hw: [L][2] = (h_l, w_l).  offsets/scales: [A,4] normalised,
 * anchors: [A,4] pixels = (offsets + scales) * [W,H,W,H]. Any output may be NULL. */
ORC_API int orc_anchors(const int32_t *level_hw, int L, int img_w, int img_h,
                        float *offsets, float *scales, float *anchors)
{
    int64_t a = 0;
    const float fw = (float)img_w, fh = (float)img_h;   /* int64 tensor promoted to f32, ref :134-136,140 */
    for (int l = 0; l < L; ++l) {
        int h = level_hw[2 * l], w = level_hw[2 * l + 1];
        if (h <= 0 || w <= 0) return -1;
        double y_min = 1.0 / h / 2.0, x_min = 1.0 / w / 2.0;          /* ref :88 (python doubles) */
        float *ys = (float *)malloc(sizeof(float) * (size_t)h);
        float *xs = (float *)malloc(sizeof(float) * (size_t)w);
        orc_linspace(y_min, 1.0 - y_min, h, ys);                       /* ref :89 */
        orc_linspace(x_min, 1.0 - x_min, w, xs);                       /* ref :90 */
        const float sc[4] = {(float)-x_min, (float)-y_min, (float)x_min, (float)y_min};   /* ref :95 */
        for (int i = 0; i < h; ++i)
            for (int j = 0; j < w; ++j, ++a) {                         /* row-major, ref :91-94 */
                const float of[4] = {xs[j], ys[i], xs[j], ys[i]};
                const float sz[4] = {fw, fh, fw, fh};
                for (int c = 0; c < 4; ++c) {
                    if (offsets) offsets[4 * a + c] = of[c];
                    if (scales)  scales[4 * a + c] = sc[c];
                    if (anchors) anchors[4 * a + c] = (of[c] + sc[c]) * sz[c];
                }
            }
        free(ys); free(xs);
    }
    return 0;
}

/* ------------------------------------------------------------------ */
/* Complete-IoU, matrix form.  tv:ops/boxes.py:404-434 (complete_box_iou), */
/* :462-480 (_box_diou_iou), :308-341 (_box_inter_union), :273-303 (box_area). */
/* Followed by .clamp(0) at ref object_detection.py:263.               */
/* ------------------------------------------------------------------ */
static inline float orc_ciou_pair(const float *a, const float *g, float atan_a, float atan_g)
{
    const float eps = 1e-7f;
    float area1 = (a[2] - a[0]) * (a[3] - a[1]);
    float area2 = (g[2] - g[0]) * (g[3] - g[1]);
    float ltx = fmaxf(a[0], g[0]), lty = fmaxf(a[1], g[1]);
    float rbx = fminf(a[2], g[2]), rby = fminf(a[3], g[3]);
    float w = rbx - ltx; if (w < 0.f) w = 0.f;           /* clamp(min=0) */
    float h = rby - lty; if (h < 0.f) h = 0.f;
    float inter = w * h;
    float uni = (area1 + area2) - inter;
    float iou = inter / uni;
    float lix = fminf(a[0], g[0]), liy = fminf(a[1], g[1]);
    float rix = fmaxf(a[2], g[2]), riy = fmaxf(a[3], g[3]);
    float wi = rix - lix; if (wi < 0.f) wi = 0.f;
    float hi = riy - liy; if (hi < 0.f) hi = 0.f;
    float diag = ((wi * wi) + (hi * hi)) + eps;
    float xp = (a[0] + a[2]) / 2.f, yp = (a[1] + a[3]) / 2.f;
    float xg = (g[0] + g[2]) / 2.f, yg = (g[1] + g[3]) / 2.f;
    float dx = xp - xg, dy = yp - yg;
    float cd = (dx * dx) + (dy * dy);
    float diou = iou - (cd / diag);
    float da = atan_a - atan_g;
    float v = (float)(4.0 / (ORC_PI * ORC_PI)) * (da * da);
    float alpha = v / (((1.f - iou) + v) + eps);
    return diou - (alpha * v);
}

static inline float orc_box_atan(const float *b)
{
    return atanf((b[2] - b[0]) / (b[3] - b[1]));
}

/* ------------------------------------------------------------------ */
/* a3: bbox_matching.  ref: object_detection.py:252-284.               */
/* Canonical form: torch.topk fills the slots of a gt */
/* that has fewer than k positive-CIoU anchors with arbitrary zero-IoU */
/* anchors; those anchors get rel_iou == 0 and an implementation-      */
/* defined assignment.  The canonical output writes assignment = -1    */
/* wherever the returned iou (rel or absolute) is not > 0; tests       */
/* canonicalise the reference's output the same way.  Exact ties       */
/* between positive values at the k-th boundary: lowest anchor index   */
/* wins.  Per-anchor ties between gts: lowest gt index (torch.max).    */
/* Precondition: no NaN CIoU (i.e. no gt with w == h == 0).            */
/* ------------------------------------------------------------------ */
ORC_API int orc_bbox_matching(const float *anchors, int64_t A, const float *gt, int64_t G,
                              int topk, int relative,
                              int64_t *assignment, float *out_iou,
                              float *best_iou /* [G] or NULL */)
{
    for (int64_t a = 0; a < A; ++a) { assignment[a] = -1; out_iou[a] = 0.f; }   /* ref :258-261 */
    if (G == 0) return 0;
    if (topk <= 0 || topk > 64 || A < topk) return -1;                          /* torch.topk raises for A < k */

    float *atan_a = (float *)malloc(sizeof(float) * (size_t)A);
    for (int64_t a = 0; a < A; ++a) atan_a[a] = orc_box_atan(anchors + 4 * a);
    float *max_v = (float *)calloc((size_t)A, sizeof(float));       /* per-anchor max over selecting gts */
    int64_t *max_g = (int64_t *)malloc(sizeof(int64_t) * (size_t)A);
    for (int64_t a = 0; a < A; ++a) max_g[a] = -1;
    float *best = (float *)calloc((size_t)G, sizeof(float));

    for (int64_t g = 0; g < G; ++g) {                       /* ascending g => first-g tie rule of torch.max */
        const float *gb = gt + 4 * g;
        float atan_g = orc_box_atan(gb);
        float tv[64]; int64_t ti[64]; int n = 0;            /* top-k list, (value desc, index asc) */
        for (int64_t a = 0; a < A; ++a) {
            float v = orc_ciou_pair(anchors + 4 * a, gb, atan_a[a], atan_g);
            if (!(v > 0.f)) continue;                       /* clamp(0): non-positive entries tie at 0 */
            if (n == topk && !(v > tv[n - 1])) continue;    /* equal value, higher index: loses */
            int p = (n < topk) ? n++ : topk - 1;
            while (p > 0 && v > tv[p - 1]) { tv[p] = tv[p - 1]; ti[p] = ti[p - 1]; --p; }
            tv[p] = v; ti[p] = a;
        }
        best[g] = n ? tv[0] : 0.f;                          /* ref :277 topk_ious[0] */
        for (int s = 0; s < n; ++s) {                       /* ref :267-273 restricted to positive slots */
            int64_t a = ti[s];
            if (tv[s] > max_v[a]) { max_v[a] = tv[s]; max_g[a] = g; }
        }
    }
    for (int64_t a = 0; a < A; ++a) {
        if (max_g[a] < 0) continue;
        assignment[a] = max_g[a];
        out_iou[a] = relative ? max_v[a] / best[max_g[a]] : max_v[a];          /* ref :279-281 */
    }
    if (best_iou) memcpy(best_iou, best, sizeof(float) * (size_t)G);
    free(atan_a); free(max_v); free(max_g); free(best);
    return 0;
}

/* ------------------------------------------------------------------ */
/* N1: QuadrilateralDetection.bbox_matching, the un-clamped variant.   */
/* ref: src/sihl/heads/quadrilateral_detection.py:266-294.             */
/* Every gt selects exactly k anchors by raw CIoU (:277-278; ties:     */
/* lowest anchor index).  Per anchor, ref :283 takes the max over gts  */
/* of iou * is_topk_match: the selected values and one zero per gt     */
/* that did not select it.  Canonical form as above: assignment = -1   */
/* wherever rel_iou is not > 0 (there the reference holds the index of */
/* an arbitrary zero entry; it only reads assignment[rel_iou > 0],     */
/* :188,:201); iou and rel_iou are 0 there exactly as in the reference.*/
/* ------------------------------------------------------------------ */
static inline float orc_nan_to_num0(float x)
{
    if (x != x) return 0.f;
    if (isinf(x)) return x > 0 ? FLT_MAX : -FLT_MAX;
    return x;
}

ORC_API int orc_quad_matching(const float *anchors, int64_t A, const float *gt, int64_t G, int topk,
                              int64_t *assignment, uint8_t *o2o, float *out_iou, float *out_rel)
{
    for (int64_t a = 0; a < A; ++a) { assignment[a] = -1; o2o[a] = 0; out_iou[a] = 0.f; out_rel[a] = 0.f; }   /* :271-276 */
    if (G == 0) return 0;
    if (topk <= 0 || topk > 64 || A < topk) return -1;
    float *atan_a = (float *)malloc(sizeof(float) * (size_t)A);
    for (int64_t a = 0; a < A; ++a) atan_a[a] = orc_box_atan(anchors + 4 * a);
    float *max_v = (float *)malloc(sizeof(float) * (size_t)A);      /* max over the selecting gts */
    int64_t *max_g = (int64_t *)malloc(sizeof(int64_t) * (size_t)A);
    int64_t *n_sel = (int64_t *)calloc((size_t)A, sizeof(int64_t));
    float *best = (float *)calloc((size_t)G, sizeof(float));
    for (int64_t a = 0; a < A; ++a) max_g[a] = -1;
    for (int64_t g = 0; g < G; ++g) {
        const float *gb = gt + 4 * g;
        float atan_g = orc_box_atan(gb);
        float tv[64]; int64_t ti[64]; int n = 0;
        for (int64_t a = 0; a < A; ++a) {
            float v = orc_ciou_pair(anchors + 4 * a, gb, atan_a[a], atan_g) + 0.f;   /* -0 -> +0 */
            if (v != v) continue;                            /* outside the defined domain */
            if (n == topk && !(v > tv[n - 1])) continue;     /* equal value, higher index: loses */
            int p = (n < topk) ? n++ : topk - 1;
            while (p > 0 && v > tv[p - 1]) { tv[p] = tv[p - 1]; ti[p] = ti[p - 1]; --p; }
            tv[p] = v; ti[p] = a;
        }
        best[g] = n ? tv[0] : 0.f;                           /* :289 topk_ious[0] */
        if (n) o2o[ti[0]] = 1;                               /* :279-280, :286 */
        for (int s2 = 0; s2 < n; ++s2) {
            int64_t a = ti[s2];
            ++n_sel[a];
            if (max_g[a] < 0 || tv[s2] > max_v[a]) { max_v[a] = tv[s2]; max_g[a] = g; }   /* ascending g: lowest gt on ties */
        }
    }
    for (int64_t a = 0; a < A; ++a) {
        if (max_g[a] < 0) continue;
        if (max_v[a] > 0.f || n_sel[a] == G) {               /* otherwise an unselected gt's zero is the max */
            out_iou[a] = max_v[a];
            out_rel[a] = orc_nan_to_num0(max_v[a] / best[max_g[a]]);                 /* :290-293 */
            if (out_rel[a] > 0.f) assignment[a] = max_g[a];
        }
    }
    free(atan_a); free(max_v); free(max_g); free(n_sel); free(best);
    return 0;
}

/* a4: the per-image loop + stack, ref :143-148.  gt in CSR form. */
ORC_API int orc_assign_batch(const float *anchors, int64_t A, const float *gt_boxes,
                             const int32_t *gt_offsets, int B, int topk, int relative,
                             int64_t *assignment /*[B,A]*/, float *out_iou /*[B,A]*/, float *best_iou /*[sumG] or NULL*/)
{
    for (int b = 0; b < B; ++b) {
        int g0 = gt_offsets[b], g1 = gt_offsets[b + 1];
        int rc = orc_bbox_matching(anchors, A, gt_boxes + 4 * (int64_t)g0, g1 - g0, topk, relative,
                                   assignment + (int64_t)b * A, out_iou + (int64_t)b * A,
                                   best_iou ? best_iou + g0 : NULL);
        if (rc) return rc;
    }
    return 0;
}

/* ------------------------------------------------------------------ */
/* a5/a6: dense losses.  ref :157-163 (BCE-with-logits vs rel==1) and  */
/* :175-180 (MSE vs rel).  sums: [0]=sum bce [1]=#(rel==1) [2]=sum mse */
/* [3]=sum rel  [6]=#(rel>0).                                          */
/* BCE-with-logits as ATen computes it (aten/src/ATen/native/Loss.cpp: */
/* (1-t)*x + max(-x,0) + log1p(exp(-|x|))).                            */
/* ------------------------------------------------------------------ */
ORC_API int orc_dense_loss(const float *loc, const float *iou_pred, const float *rel, int64_t n,
                           double *sums /* [8], accumulated into */)
{
    double bce = 0, n1 = 0, mse = 0, rs = 0, np = 0;
    for (int64_t i = 0; i < n; ++i) {
        float x = loc[i], r = rel[i];
        float t = (r == 1.0f) ? 1.f : 0.f;
        float l = (1.f - t) * x + fmaxf(-x, 0.f) + log1pf(expf(-fabsf(x)));
        bce += l; n1 += t;
        if (iou_pred) { float d = iou_pred[i] - r; mse += (double)(d * d); }
        rs += r;
        if (r > 0.f) np += 1;
    }
    sums[0] += bce; sums[1] += n1; sums[2] += mse; sums[3] += rs; sums[6] += np;
    return 0;
}

/* ------------------------------------------------------------------ */
/* a8: box decode + CIoU loss on one row.  ref :189-197;               */
/* tv:ops/ciou_loss.py:47-64, tv:ops/diou_loss.py:64-91,               */
/* tv:ops/_utils.py:87-106.  pred = offsets + scales*exp(raw)          */
/* (normalised), target = gt_px / [W,H,W,H].                           */
/* ------------------------------------------------------------------ */
ORC_API float orc_ciou_loss_row(const float *p, const float *t)
{
    const float eps = 1e-7f;
    float x1 = p[0], y1 = p[1], x2 = p[2], y2 = p[3];
    float x1g = t[0], y1g = t[1], x2g = t[2], y2g = t[3];
    float xk1 = fmaxf(x1, x1g), yk1 = fmaxf(y1, y1g);
    float xk2 = fminf(x2, x2g), yk2 = fminf(y2, y2g);
    float inter = 0.f;
    if (yk2 > yk1 && xk2 > xk1) inter = (xk2 - xk1) * (yk2 - yk1);
    float uni = (((x2 - x1) * (y2 - y1)) + ((x2g - x1g) * (y2g - y1g))) - inter;
    float iou = inter / (uni + eps);
    float xc1 = fminf(x1, x1g), yc1 = fminf(y1, y1g);
    float xc2 = fmaxf(x2, x2g), yc2 = fmaxf(y2, y2g);
    float diag = (((xc2 - xc1) * (xc2 - xc1)) + ((yc2 - yc1) * (yc2 - yc1))) + eps;
    float xp = (x2 + x1) / 2.f, yp = (y2 + y1) / 2.f;
    float xg = (x1g + x2g) / 2.f, yg = (y1g + y2g) / 2.f;
    float cd = ((xp - xg) * (xp - xg)) + ((yp - yg) * (yp - yg));
    float loss = (1.f - iou) + (cd / diag);
    float wp = x2 - x1, hp = y2 - y1, wg = x2g - x1g, hg = y2g - y1g;
    float da = atanf(wg / hg) - atanf(wp / hp);
    float v = (float)(4.0 / (ORC_PI * ORC_PI)) * (da * da);
    float alpha = v / (((1.f - iou) + v) + eps);
    return loss + alpha * v;
}

/* a7-a9: positives.  Rows are visited in row-major (b, a) order over rel>0,
 * exactly the order of flat_feats[o2m_mask] (ref :182-184).
 * box_raw / cls_logits are addressed through row_of[p]: pass dense_rows=1 for
 * dense maps [B*A, 4] / [B*A, C] (row = b*A+a) or 0 for compact [P, ...] rows.
 * Outputs (optional): pos_index[P] = b*A+a.  sums[4] += sum w*ciou_loss,
 * sums[5] += sum w*ce.  Returns P. */
ORC_API int64_t orc_pos_loss(const float *rel, const int64_t *assignment, int B, int64_t A,
                             const float *offsets, const float *scales, int img_w, int img_h,
                             const float *gt_boxes, const int64_t *gt_classes, const int32_t *gt_offsets,
                             const float *box_raw, const float *cls_logits, int C, int dense_rows,
                             double *sums, int32_t *pos_index,
                             float *box_loss_rows /* [P] or NULL */, float *ce_rows /* [P] or NULL */)
{
    int64_t p = 0;
    const float size[4] = {(float)img_w, (float)img_h, (float)img_w, (float)img_h};
    for (int b = 0; b < B; ++b)
        for (int64_t a = 0; a < A; ++a) {
            int64_t flat = (int64_t)b * A + a;
            float w = rel[flat];
            if (!(w > 0.f)) continue;
            int64_t row = dense_rows ? flat : p;
            int64_t g = gt_offsets[b] + assignment[flat];
            if (box_raw) {
                float pred[4], tgt[4];
                for (int c = 0; c < 4; ++c) {
                    pred[c] = offsets[4 * a + c] + scales[4 * a + c] * expf(box_raw[4 * row + c]);   /* ref :189 */
                    tgt[c] = gt_boxes[4 * g + c] / size[c];                                        /* ref :195 */
                }
                float l = orc_ciou_loss_row(pred, tgt);
                sums[4] += (double)(w * l);                                                        /* ref :197 */
                if (box_loss_rows) box_loss_rows[p] = l;
            }
            if (cls_logits) {                                                                      /* ref :205-208 */
                const float *z = cls_logits + (int64_t)C * row;
                float m = z[0];
                for (int c = 1; c < C; ++c) m = fmaxf(m, z[c]);
                double se = 0;
                for (int c = 0; c < C; ++c) se += exp((double)(z[c] - m));
                float ce = (float)(log(se) + (double)m - (double)z[gt_classes[g]]);
                sums[5] += (double)(w * ce);
                if (ce_rows) ce_rows[p] = ce;
            }
            if (pos_index) pos_index[p] = (int32_t)flat;
            ++p;
        }
    return p;
}

/* a10: ref :163-172, :180, :197, :208, :210.  losses = [location, box, class, iou, total]. */
ORC_API void orc_loss_finalize(const double *sums, float *losses)
{
    double loc = sums[0] / sums[1];
    if (sums[6] == 0) {          /* rel_iou.max() == 0 early-out */
        losses[0] = (float)loc; losses[1] = losses[2] = losses[3] = 0.f; losses[4] = (float)loc;
        return;
    }
    double iou = sums[2] / sums[3], box = sums[4] / sums[3], cls = sums[5] / sums[3];
    losses[0] = (float)loc; losses[1] = (float)box; losses[2] = (float)cls; losses[3] = (float)iou;
    losses[4] = (float)(loc + 10.0 * box + cls + iou);
}

/* ------------------------------------------------------------------ */
/* a11: forward tail.  ref :108-121.  top-K of the location logits per */
/* image, sorted by (logit desc, index asc); torch.topk's order among  */
/* exactly equal logits is implementation-defined.  */
/* ------------------------------------------------------------------ */
typedef struct { float v; int64_t i; } orc_vi;
static int orc_cmp_desc(const void *x, const void *y)
{
    const orc_vi *a = (const orc_vi *)x, *b = (const orc_vi *)y;
    if (a->v > b->v) return -1;
    if (a->v < b->v) return 1;
    return (a->i < b->i) ? -1 : (a->i > b->i);
}

ORC_API int orc_topk_rows(const float *loc, int B, int64_t A, int K,
                          int64_t *idx /*[B,K]*/, float *top_logits /*[B,K]*/)
{
    if (K > A) return -1;
    orc_vi *buf = (orc_vi *)malloc(sizeof(orc_vi) * (size_t)A);
    for (int b = 0; b < B; ++b) {
        for (int64_t a = 0; a < A; ++a) { buf[a].v = loc[(int64_t)b * A + a]; buf[a].i = a; }
        qsort(buf, (size_t)A, sizeof(orc_vi), orc_cmp_desc);
        for (int k = 0; k < K; ++k) { idx[(int64_t)b * K + k] = buf[k].i; top_logits[(int64_t)b * K + k] = buf[k].v; }
    }
    free(buf);
    return 0;
}

/* sigmoid as ATen: 1 / (1 + exp(-x)) in fp32 */
static inline float orc_sigmoid(float x) { return 1.f / (1.f + expf(-x)); }

/* ref :113-121 on gathered rows.  cls_rows [B,K,C], box_rows [B,K,4]. */
ORC_API int orc_decode_rows(const float *top_logits, const int64_t *idx, int B, int K,
                            const float *cls_rows, int C, const float *box_rows,
                            const float *offsets, const float *scales, int img_w, int img_h,
                            int64_t *num_instances, float *scores, int64_t *classes, float *boxes)
{
    const float size[4] = {(float)img_w, (float)img_h, (float)img_w, (float)img_h};
    for (int b = 0; b < B; ++b) {
        int64_t n = 0;
        for (int k = 0; k < K; ++k) {
            int64_t r = (int64_t)b * K + k;
            float s = orc_sigmoid(top_logits[r]);
            scores[r] = s;
            if (s > 0.5f) ++n;                                   /* ref :114 */
            const float *z = cls_rows + (int64_t)C * r;
            int best = 0;
            for (int c = 1; c < C; ++c) if (z[c] > z[best]) best = c;   /* first max, ref :117 */
            classes[r] = best;
            int64_t a = idx[r];
            for (int c = 0; c < 4; ++c)
                boxes[4 * r + c] = (offsets[4 * a + c] + scales[4 * a + c] * expf(box_rows[4 * r + c])) * size[c];  /* ref :121 */
        }
        num_instances[b] = n;
    }
    return 0;
}

/* ------------------------------------------------------------------ */
/* a15 (extension; no reference code): class-aware greedy NMS with the */
/* semantics of torchvision's per-class path, tv:ops/boxes.py:102-120  */
/* (_batched_nms_vanilla) over the CPU kernel                          */
/* (torchvision/csrc/ops/cpu/nms_kernel.cpp): visit boxes by           */
/* (score desc, index asc); a kept box suppresses every later box of   */
/* the same class with inter/(area_i+area_j-inter) > thr.              */
/* keep: indices in (score desc, index asc) order.  Returns M.         */
/* ------------------------------------------------------------------ */
ORC_API int64_t orc_batched_nms(const float *boxes, const float *scores, const int64_t *classes,
                                int64_t N, float iou_thr, int64_t *keep)
{
    if (N == 0) return 0;
    orc_vi *ord = (orc_vi *)malloc(sizeof(orc_vi) * (size_t)N);
    for (int64_t i = 0; i < N; ++i) { ord[i].v = scores[i]; ord[i].i = i; }
    qsort(ord, (size_t)N, sizeof(orc_vi), orc_cmp_desc);
    unsigned char *dead = (unsigned char *)calloc((size_t)N, 1);
    int64_t m = 0;
    for (int64_t oi = 0; oi < N; ++oi) {
        int64_t i = ord[oi].i;
        if (dead[i]) continue;
        keep[m++] = i;
        const float *bi = boxes + 4 * i;
        float iarea = (bi[2] - bi[0]) * (bi[3] - bi[1]);
        for (int64_t oj = oi + 1; oj < N; ++oj) {
            int64_t j = ord[oj].i;
            if (dead[j] || classes[j] != classes[i]) continue;
            const float *bj = boxes + 4 * j;
            float xx1 = fmaxf(bi[0], bj[0]), yy1 = fmaxf(bi[1], bj[1]);
            float xx2 = fminf(bi[2], bj[2]), yy2 = fminf(bi[3], bj[3]);
            float w = fmaxf(0.f, xx2 - xx1), h = fmaxf(0.f, yy2 - yy1);
            float inter = w * h;
            float jarea = (bj[2] - bj[0]) * (bj[3] - bj[1]);
            float ovr = inter / ((iarea + jarea) - inter);
            if (ovr > iou_thr) dead[j] = 1;
        }
    }
    free(ord); free(dead);
    return m;
}

/* Dense postprocess (extension): every location is decoded — class = first
 * argmax of its C class logits, score = sigmoid(loc logit) (the score/class
 * semantics of ref forward :113,:117), candidates = score > score_thr, boxes
 * decoded as ref :121 — then class-aware NMS per image and the first K kept
 * detections in (score desc, location asc) order; the rest is zero-padded.
 * num_instances = number of valid (unpadded) detections. */
ORC_API int orc_dense_postprocess(const float *loc, const float *cls, const float *box_raw,
                                  int B, int64_t A, int C,
                                  const float *offsets, const float *scales, int img_w, int img_h,
                                  float score_thr, float iou_thr, int K,
                                  int64_t *num_instances, float *scores, int64_t *classes, float *boxes,
                                  int64_t *n_candidates /* [B] or NULL */)
{
    const float size[4] = {(float)img_w, (float)img_h, (float)img_w, (float)img_h};
    float *cb = (float *)malloc(sizeof(float) * 4 * (size_t)A);
    float *cs = (float *)malloc(sizeof(float) * (size_t)A);
    int64_t *cc = (int64_t *)malloc(sizeof(int64_t) * (size_t)A);
    int64_t *keep = (int64_t *)malloc(sizeof(int64_t) * (size_t)A);
    for (int b = 0; b < B; ++b) {
        int64_t n = 0;
        for (int64_t a = 0; a < A; ++a) {
            int64_t flat = (int64_t)b * A + a;
            float s = orc_sigmoid(loc[flat]);
            if (!(s > score_thr)) continue;
            const float *z = cls + (int64_t)C * flat;
            int best = 0;
            for (int c = 1; c < C; ++c) if (z[c] > z[best]) best = c;
            for (int c = 0; c < 4; ++c)
                cb[4 * n + c] = (offsets[4 * a + c] + scales[4 * a + c] * expf(box_raw[4 * flat + c])) * size[c];
            cs[n] = s; cc[n] = best; ++n;
        }
        if (n_candidates) n_candidates[b] = n;
        int64_t m = orc_batched_nms(cb, cs, cc, n, iou_thr, keep);
        if (m > K) m = K;
        num_instances[b] = m;
        for (int k = 0; k < K; ++k) {
            int64_t r = (int64_t)b * K + k;
            if (k < m) {
                int64_t j = keep[k];
                scores[r] = cs[j]; classes[r] = cc[j];
                for (int c = 0; c < 4; ++c) boxes[4 * r + c] = cb[4 * j + c];
            } else {
                scores[r] = 0.f; classes[r] = 0;
                for (int c = 0; c < 4; ++c) boxes[4 * r + c] = 0.f;
            }
        }
    }
    free(cb); free(cs); free(cc); free(keep);
    return 0;
}

ORC_API int orc_version(void) { return 1; }

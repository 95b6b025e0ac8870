"""CPU restatement of the COCO detection evaluation behind the reference's validation mAP.

TEST INFRASTRUCTURE ONLY (only ``tests/`` import this; nothing under ``sihl_b200/`` does).

**PARITY UNPINNED.**  The reference calls ``torchmetrics.detection.mean_ap.MeanAveragePrecision(
max_detection_thresholds=[1, 10, K], backend="faster_coco_eval")`` (ref src/sihl/heads/object_detection.py:219-237,
:245).  The arithmetic lives in third-party dependencies that are NOT in ``/root/reference`` and NOT installed here:
``torchmetrics==1.6.1`` and ``faster-coco-eval==1.6.5`` (``requirements.lock:21,129``), the latter a C++ port of
``pycocotools.cocoeval.COCOeval``.  This file restates that published algorithm — ``COCOeval.evaluateImg`` (greedy
matching), ``COCOeval.accumulate`` (precision / recall tables) and ``COCOeval.summarize`` with torchmetrics' key names —
in plain Python loops over numpy arrays.  The reference's own tests hold no value for this step
(``tests/heads/test_object_detection.py:72-80`` only asserts that the metrics dict is non-empty), and neither library can be
run here to generate fixtures; the anchors are (a) the ONE published input/output vector of the real library, the example
in ``MeanAveragePrecision``'s docstring (``TORCHMETRICS_DOC_EXAMPLE`` below: all 12 summary values), and (b) known-answer
cases in closed form (perfect detections, greedy order across IoU thresholds, a missed object + a false positive, the
area-range ignore rules), all in ``tests/test_map_oracle.py``.  One published vector is not a pin: the label stays.

Conventions restated from torchmetrics' ``_get_coco_format`` / faster_coco_eval: boxes xyxy -> xywh with w, h computed in
fp32; areas = w * h in double; iscrowd = 0; detections of an image sorted by score descending with a stable sort;
ground truth sorted "ignored last" with a stable sort; ``maxDets`` slicing per (image, category).
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np

IOU_THRESHOLDS = np.linspace(0.5, 0.95, 10)
RECALL_THRESHOLDS = np.linspace(0.0, 1.0, 101)
AREA_RANGES = ((0.0, 1e5 ** 2), (0.0, 32.0 ** 2), (32.0 ** 2, 96.0 ** 2), (96.0 ** 2, 1e5 ** 2))   # all, small, medium, large


# The one published input/output vector of the real library this path replaces: the example in the docstring of
# ``torchmetrics.detection.mean_ap.MeanAveragePrecision`` (torchmetrics 1.6.x, the version the reference pins in
# requirements.lock) — one detection, one ground-truth box, IoU 304 / 392 = 0.7755, so 6 of the 10 IoU thresholds match.
TORCHMETRICS_DOC_EXAMPLE = dict(
    det_boxes=[[258.0, 41.0, 606.0, 285.0]], det_scores=[0.536], det_classes=[0],
    gt_boxes=[[214.0, 41.0, 562.0, 285.0]], gt_classes=[0],
    want={"map": 0.6, "map_50": 1.0, "map_75": 1.0, "map_small": -1.0, "map_medium": -1.0, "map_large": 0.6,
          "mar_1": 0.6, "mar_10": 0.6, "mar_100": 0.6, "mar_small": -1.0, "mar_medium": -1.0, "mar_large": 0.6})


def _wh32(b: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    b = np.asarray(b, np.float32).reshape(-1, 4)
    return (b[:, 2] - b[:, 0]).astype(np.float32).astype(np.float64), (b[:, 3] - b[:, 1]).astype(np.float32).astype(np.float64)


def box_area(b: np.ndarray) -> np.ndarray:
    w, h = _wh32(b)
    return w * h


def box_iou(d: np.ndarray, g: np.ndarray) -> np.ndarray:
    """maskUtils.iou for boxes (bbIou, iscrowd = 0), double precision -> [D, G]."""
    d = np.asarray(d, np.float32).reshape(-1, 4); g = np.asarray(g, np.float32).reshape(-1, 4)
    dw, dh = _wh32(d); gw, gh = _wh32(g)
    dx, dy, gx, gy = d[:, 0].astype(np.float64), d[:, 1].astype(np.float64), g[:, 0].astype(np.float64), g[:, 1].astype(np.float64)
    out = np.zeros((len(d), len(g)))
    for i in range(len(d)):
        for j in range(len(g)):
            w = min(dx[i] + dw[i], gx[j] + gw[j]) - max(dx[i], gx[j])
            if w <= 0:
                continue
            h = min(dy[i] + dh[i], gy[j] + gh[j]) - max(dy[i], gy[j])
            if h <= 0:
                continue
            inter = w * h
            out[i, j] = inter / (dw[i] * dh[i] + gw[j] * gh[j] - inter)
    return out


def match_image(det_boxes, det_scores, det_classes, gt_boxes, gt_classes, iou_thresholds=IOU_THRESHOLDS,
                area_ranges=AREA_RANGES):
    """COCOeval.evaluateImg for every category / area range of ONE image, all detections (maxDet = K).
    Returns dict(order [K] rank -> detection, dt_match [NA,T,K] by rank: gt index or -1, dt_ignore [NA,T,K],
    gt_ignore [NA,G])."""
    det_scores = np.asarray(det_scores, np.float32)
    K, G, T, NA = len(det_scores), len(gt_boxes), len(iou_thresholds), len(area_ranges)
    order = np.argsort(-det_scores.astype(np.float64), kind="mergesort")
    ious = box_iou(det_boxes, gt_boxes) if G else np.zeros((K, 0))
    d_area, g_area = box_area(det_boxes), box_area(gt_boxes) if G else np.zeros(0)
    dt_match = -np.ones((NA, T, K), np.int32)
    dt_ignore = np.zeros((NA, T, K), np.uint8)
    gt_ignore = np.zeros((NA, G), np.uint8)
    for a, (lo, hi) in enumerate(area_ranges):
        gt_ignore[a] = (g_area < lo) | (g_area > hi)
        for t, thr in enumerate(iou_thresholds):
            taken = np.zeros(G, bool)
            for r, d in enumerate(order):
                gts = [g for g in range(G) if gt_classes[g] == det_classes[d]]
                gts = sorted(gts, key=lambda g: gt_ignore[a, g])                   # stable: ignored last
                best, m = min(thr, 1 - 1e-10), -1
                for g in gts:
                    if taken[g]:
                        continue
                    if m > -1 and gt_ignore[a, m] == 0 and gt_ignore[a, g] == 1:
                        break
                    if ious[d, g] < best:
                        continue
                    best, m = ious[d, g], g
                if m == -1:
                    dt_ignore[a, t, r] = d_area[d] < lo or d_area[d] > hi
                    continue
                dt_ignore[a, t, r] = gt_ignore[a, m]
                dt_match[a, t, r] = m
                taken[m] = True
    return dict(order=order.astype(np.int32), dt_match=dt_match, dt_ignore=dt_ignore, gt_ignore=gt_ignore)


def accumulate(images: List[dict], num_classes_seen: Sequence[int], max_dets: Sequence[int] = (1, 10, 100),
               iou_thresholds=IOU_THRESHOLDS, area_ranges=AREA_RANGES):
    """COCOeval.accumulate.  ``images``: per image dict(scores [K] by rank, classes [K] by rank, dt_match, dt_ignore by
    rank, gt_classes [G], gt_ignore [NA,G]).  Returns (precision [T,R,Kc,NA,M], recall [T,Kc,NA,M]), -1 where undefined."""
    T, R, NA, M = len(iou_thresholds), len(RECALL_THRESHOLDS), len(area_ranges), len(max_dets)
    cats = list(num_classes_seen)
    precision = -np.ones((T, R, len(cats), NA, M))
    recall = -np.ones((T, len(cats), NA, M))
    for k, cat in enumerate(cats):
        for a in range(NA):
            for mi, max_det in enumerate(max_dets):
                scores, dtm, dtig, npig = [], [], [], 0
                for im in images:
                    sel = np.nonzero(im["classes"] == cat)[0][:max_det]            # ranks of this category, best first
                    gsel = im["gt_classes"] == cat
                    if len(sel) == 0 and not gsel.any():
                        continue
                    scores.append(im["scores"][sel])
                    dtm.append(im["dt_match"][a][:, sel] >= 0)
                    dtig.append(im["dt_ignore"][a][:, sel] != 0)
                    npig += int((im["gt_ignore"][a][gsel] == 0).sum())
                if not scores or npig == 0:
                    continue
                scores = np.concatenate(scores)
                inds = np.argsort(-scores.astype(np.float64), kind="mergesort")
                dtm = np.concatenate(dtm, axis=1)[:, inds]
                dtig = np.concatenate(dtig, axis=1)[:, inds]
                tps = np.logical_and(dtm, np.logical_not(dtig))
                fps = np.logical_and(np.logical_not(dtm), np.logical_not(dtig))
                tp_sum = np.cumsum(tps, axis=1).astype(np.float64)
                fp_sum = np.cumsum(fps, axis=1).astype(np.float64)
                for t in range(T):
                    tp, fp = tp_sum[t], fp_sum[t]
                    nd = len(tp)
                    rc = tp / npig
                    pr = tp / (fp + tp + np.spacing(1))
                    recall[t, k, a, mi] = rc[-1] if nd else 0
                    pr = pr.tolist()
                    for i in range(nd - 1, 0, -1):
                        if pr[i] > pr[i - 1]:
                            pr[i - 1] = pr[i]
                    q = np.zeros(R)
                    idx = np.searchsorted(rc, RECALL_THRESHOLDS, side="left")
                    for ri, pi in enumerate(idx):
                        if pi < nd:
                            q[ri] = pr[pi]
                    precision[t, :, k, a, mi] = q
    return precision, recall


def summarize(precision: np.ndarray, recall: np.ndarray, max_dets: Sequence[int] = (1, 10, 100),
              iou_thresholds=IOU_THRESHOLDS) -> Dict[str, float]:
    """COCOeval.summarize with torchmetrics' key names (map, map_50, ..., mar_<maxDet>, mar_small, ...)."""
    def ap(iou=None, area=0, m=len(max_dets) - 1):
        s = precision[:, :, :, area, m]
        if iou is not None:
            s = s[[i for i, t in enumerate(iou_thresholds) if abs(t - iou) < 1e-9]]
        s = s[s > -1]
        return float(s.mean()) if s.size else -1.0

    def ar(area=0, m=len(max_dets) - 1):
        s = recall[:, :, area, m]
        s = s[s > -1]
        return float(s.mean()) if s.size else -1.0

    out = {"map": ap(), "map_50": ap(0.5), "map_75": ap(0.75), "map_small": ap(area=1), "map_medium": ap(area=2),
           "map_large": ap(area=3)}
    for mi, md in enumerate(max_dets):
        out[f"mar_{md}"] = ar(m=mi)
    out.update({"mar_small": ar(area=1), "mar_medium": ar(area=2), "mar_large": ar(area=3)})
    return out

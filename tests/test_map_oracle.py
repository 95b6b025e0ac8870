"""Known-answer tests of the COCO-evaluation oracle (oracle/map_oracle.py) and of the host half of
``sihl_b200.metrics.DetectionMAP`` (accumulate / summarize), without a GPU.  torchmetrics / faster_coco_eval are not
installed here, so these are the anchors the restatement has: cases whose COCO metrics are known in closed form."""
import numpy as np
import pytest
import torch

from oracle import map_oracle as mo
from sihl_b200 import metrics


def _image(det_boxes, det_scores, det_classes, gt_boxes, gt_classes):
    det_boxes = np.asarray(det_boxes, np.float32).reshape(-1, 4); gt_boxes = np.asarray(gt_boxes, np.float32).reshape(-1, 4)
    det_scores = np.asarray(det_scores, np.float32); det_classes = np.asarray(det_classes, np.int64)
    gt_classes = np.asarray(gt_classes, np.int64)
    m = mo.match_image(det_boxes, det_scores, det_classes, gt_boxes, gt_classes)
    order = m["order"]
    return dict(scores=det_scores[order], classes=det_classes[order], dt_match=m["dt_match"], dt_ignore=m["dt_ignore"],
                gt_classes=gt_classes, gt_ignore=m["gt_ignore"]), m


def _summary(images, cats):
    p, r = mo.accumulate(images, cats)
    return mo.summarize(p, r)


def test_perfect_detections_score_one():
    gt = [[10, 10, 110, 110], [200, 50, 400, 300]]
    im, m = _image(gt, [0.9, 0.8], [1, 2], gt, [1, 2])
    assert (m["dt_match"][0] >= 0).all()
    s = _summary([im], [1, 2])
    assert s["map"] == pytest.approx(1.0) and s["map_50"] == pytest.approx(1.0) and s["mar_100"] == pytest.approx(1.0)
    assert s["mar_1"] == pytest.approx(1.0)                       # one object per category
    assert s["map_small"] == -1.0                                 # no small ground truth at all
    assert s["map_large"] == pytest.approx(1.0)


def test_iou_thresholds_and_greedy_order():
    """One gt 100x100; detection A (score .9) has IoU 0.64, detection B (score .8) IoU 1.0.  A takes the gt at the
    thresholds <= .6 (B becomes a false positive there); at the higher ones A fails and B matches."""
    gt = [[0, 0, 100, 100]]
    im, m = _image([[0, 0, 100, 64], [0, 0, 100, 100]], [0.9, 0.8], [0, 0], gt, [0])
    dtm = m["dt_match"][0]                                        # [T, K] by rank: rank 0 = A, rank 1 = B
    assert (dtm[:3, 0] == 0).all() and (dtm[:3, 1] == -1).all()   # .50 .55 .60: A matched, B not
    assert (dtm[3:, 0] == -1).all() and (dtm[3:, 1] == 0).all()   # .65 ... .95: A unmatched, B matched
    s = _summary([im], [0])
    # per threshold: t <= .6 -> precision 1 at recall 1 (TP first): AP 1; t >= .65 -> FP then TP: precision 1/2: AP .5
    assert s["map_50"] == pytest.approx(1.0)
    assert s["map_75"] == pytest.approx(0.5, abs=1e-6)
    assert s["map"] == pytest.approx((3 * 1.0 + 7 * 0.5) / 10, abs=1e-6)


def test_missed_object_and_false_positive():
    gt = [[0, 0, 50, 50], [100, 100, 200, 200]]
    im, _ = _image([[0, 0, 50, 50], [300, 300, 350, 350]], [0.9, 0.7], [3, 3], gt, [3, 3])
    s = _summary([im], [3])
    # recall reaches 0.5 only: precision 1 for the 51 recall thresholds <= 0.5, 0 beyond
    assert s["map"] == pytest.approx(51 / 101, abs=1e-6) and s["mar_100"] == pytest.approx(0.5)


def test_area_ranges_ignore_rules():
    """A small gt (20x20) and a large one: in the 'large' range the small gt is ignored, and a detection matched to
    it is ignored rather than counted as a false positive."""
    gt = [[0, 0, 20, 20], [100, 100, 300, 300]]
    im, m = _image([[0, 0, 20, 20], [100, 100, 300, 300]], [0.9, 0.8], [0, 0], gt, [0, 0])
    assert m["gt_ignore"][3].tolist() == [1, 0] and m["gt_ignore"][1].tolist() == [0, 1]
    assert m["dt_ignore"][3][0].tolist() == [1, 0]                # large range: det 0 matched the ignored small gt
    s = _summary([im], [0])
    assert s["map_small"] == pytest.approx(1.0) and s["map_large"] == pytest.approx(1.0) and s["map_medium"] == -1.0


def test_published_torchmetrics_docstring_example():
    e = mo.TORCHMETRICS_DOC_EXAMPLE
    im, m = _image(e["det_boxes"], e["det_scores"], e["det_classes"], e["gt_boxes"], e["gt_classes"])
    assert (m["dt_match"][0][:6, 0] == 0).all() and (m["dt_match"][0][6:, 0] == -1).all()
    s = _summary([im], [0])
    for k, v in e["want"].items():
        assert s[k] == pytest.approx(v, abs=1e-6), k


def test_metrics_accumulate_equals_the_oracle_on_random_matches():
    """The vectorised host half of DetectionMAP (cumulative sums per category) == the loop restatement, fed with the
    same random match tables (several images, categories, ties in the scores, categories without gt / without dets)."""
    rng = np.random.RandomState(0)
    images = []
    K, NA, T = 12, 4, 10
    for _ in range(7):
        G = rng.randint(0, 6)
        scores = np.sort(np.round(rng.uniform(0, 1, K), 1).astype(np.float32))[::-1].copy()
        im = dict(scores=scores, classes=rng.randint(0, 4, K).astype(np.int64),
                  dt_match=np.where(rng.uniform(size=(NA, T, K)) < 0.5, rng.randint(0, max(G, 1), (NA, T, K)), -1).astype(np.int32),
                  dt_ignore=(rng.uniform(size=(NA, T, K)) < 0.2).astype(np.uint8),
                  gt_classes=rng.randint(0, 5, G).astype(np.int64), gt_ignore=(rng.uniform(size=(NA, G)) < 0.3).astype(np.uint8))
        images.append(im)
    cats = sorted(set(int(c) for im in images for c in im["classes"]) | set(int(c) for im in images for c in im["gt_classes"]))
    want = _summary(images, cats)
    m = metrics.DetectionMAP(max_detection_thresholds=[1, 10, 100], sync_dist=False)
    m._host_state = lambda: [(im["scores"], im["classes"], im["dt_match"] >= 0, im["dt_ignore"] != 0, im["gt_classes"],
                              im["gt_ignore"]) for im in images]
    got = m.compute()
    for k, v in want.items():
        assert float(got[k]) == pytest.approx(v, abs=1e-6), k

"""N3 — GPU-side matching for the validation mAP (sihl_od_map_match, sihl_b200.metrics.DetectionMAP) against the
COCOeval restatement oracle/map_oracle.py.  Parity is unpinned (torchmetrics / faster_coco_eval are absent): both sides
follow the published COCOeval algorithm; the oracle's own anchors are the known-answer cases of tests/test_map_oracle.py."""
import numpy as np
import pytest
import torch

from oracle import map_oracle as mo
from sihl_b200 import metrics, ops, synth
from sihl_b200.heads import ObjectDetection

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _t(x):
    return torch.from_numpy(np.ascontiguousarray(x)).to(DEV)


def _random_batch(seed, B=6, K=40, C=5, size=400, gmax=14):
    rng = np.random.RandomState(seed)
    counts = [0] + [int(rng.randint(1, gmax)) for _ in range(B - 1)]
    gt = synth.gt_batch_np(seed, B, size, size, C, gmax, counts=counts)
    gt_boxes = gt.boxes.copy()
    gt_boxes[::4] = np.round(gt_boxes[::4] / 24) * 24 + np.array([0, 0, 24, 24], np.float32)   # some small / exact-lattice boxes
    det_boxes = np.zeros((B, K, 4), np.float32); det_scores = np.zeros((B, K), np.float32); det_classes = np.zeros((B, K), np.int64)
    for b in range(B):
        g0, g1 = gt.offsets[b], gt.offsets[b + 1]
        for k in range(K):
            if g1 > g0 and rng.uniform() < 0.6:                      # a jittered (sometimes exact) copy of a gt
                g = rng.randint(g0, g1)
                jit = rng.choice([0.0, 2.0, 8.0, 25.0])
                det_boxes[b, k] = gt_boxes[g] + rng.uniform(-jit, jit, 4).astype(np.float32)
                det_classes[b, k] = gt.classes[g] if rng.uniform() < 0.8 else rng.randint(0, C)
            else:
                xy = rng.uniform(0, size - 60, 2)
                det_boxes[b, k] = [xy[0], xy[1], xy[0] + rng.uniform(5, 150), xy[1] + rng.uniform(5, 150)]
                det_classes[b, k] = rng.randint(0, C)
        det_scores[b] = np.round(rng.uniform(0, 1, K), 2)              # two decimals: plenty of exact ties
    det_boxes[..., 2:] = np.maximum(det_boxes[..., 2:], det_boxes[..., :2] + 1)
    return det_boxes, det_scores, det_classes, gt_boxes, gt.classes, gt.offsets


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_map_match_equals_the_cocoeval_restatement(seed):
    db, ds, dc, gb, gc_, go = _random_batch(seed)
    if seed == 3:                                                      # duplicate gts: equal IoUs -> the LAST one wins
        gb[go[2] + 1] = gb[go[2]]; gc_[go[2] + 1] = gc_[go[2]]
    res = ops.map_match(_t(db), _t(ds), _t(dc), _t(gb), _t(gc_), _t(go))
    order, dtm, dti, gti = (res[k].cpu().numpy() for k in ("det_order", "dt_match", "dt_ignore", "gt_ignore"))
    for b in range(db.shape[0]):
        g0, g1 = go[b], go[b + 1]
        want = mo.match_image(db[b], ds[b], dc[b], gb[g0:g1], gc_[g0:g1])
        np.testing.assert_array_equal(order[b], want["order"])
        got_m = np.where(dtm[b] >= 0, dtm[b] - g0, -1)
        np.testing.assert_array_equal(got_m, want["dt_match"])
        np.testing.assert_array_equal(dti[b], want["dt_ignore"])
        np.testing.assert_array_equal(gti[:, g0:g1], want["gt_ignore"])
    assert (dtm >= 0).any() and (dti != 0).any()


def test_detection_map_metric_equals_the_oracle_pipeline():
    m = metrics.DetectionMAP(max_detection_thresholds=[1, 10, 40], sync_dist=False)
    images = []
    for seed in (11, 12):
        db, ds, dc, gb, gc_, go = _random_batch(seed)
        counts = np.diff(go).tolist()
        if seed == 11:
            m.update_batch(_t(ds), _t(dc), _t(db), _t(gb), _t(gc_), counts)
        else:                                                          # torchmetrics' per-image dict signature
            m.update([{"scores": _t(ds[b]), "labels": _t(dc[b]), "boxes": _t(db[b])} for b in range(len(counts))],
                     [{"labels": _t(gc_[go[b]:go[b + 1]]), "boxes": _t(gb[go[b]:go[b + 1]])} for b in range(len(counts))])
        for b in range(len(counts)):
            g0, g1 = go[b], go[b + 1]
            w = mo.match_image(db[b], ds[b], dc[b], gb[g0:g1], gc_[g0:g1])
            images.append(dict(scores=ds[b][w["order"]], classes=dc[b][w["order"]], dt_match=w["dt_match"],
                               dt_ignore=w["dt_ignore"], gt_classes=gc_[g0:g1], gt_ignore=w["gt_ignore"]))
    cats = sorted(set(int(c) for im in images for c in im["classes"]) | set(int(c) for im in images for c in im["gt_classes"]))
    p, r = mo.accumulate(images, cats, max_dets=(1, 10, 40))
    want = mo.summarize(p, r, max_dets=(1, 10, 40))
    got = m.compute()
    assert set(want) <= set(got)
    for k, v in want.items():
        assert float(got[k]) == pytest.approx(v, abs=1e-6), k
    assert 0.0 < float(got["map"]) < 1.0 and float(got["mar_40"]) >= float(got["mar_1"])


def test_detection_map_reproduces_the_published_torchmetrics_example():
    """The docstring example of ``torchmetrics.detection.mean_ap.MeanAveragePrecision`` (the library the reference calls,
    ref :219-237) through the GPU matcher + DetectionMAP: map 0.6, map_50 1, map_75 1, mar_* 0.6, small / medium -1."""
    e = mo.TORCHMETRICS_DOC_EXAMPLE
    m = metrics.DetectionMAP(sync_dist=False)
    m.update([{"boxes": _t(np.asarray(e["det_boxes"], np.float32)), "scores": _t(np.asarray(e["det_scores"], np.float32)),
               "labels": _t(np.asarray(e["det_classes"], np.int64))}],
             [{"boxes": _t(np.asarray(e["gt_boxes"], np.float32)), "labels": _t(np.asarray(e["gt_classes"], np.int64))}])
    got = m.compute()
    for k, v in e["want"].items():
        assert float(got[k]) == pytest.approx(v, abs=1e-6), k


def test_head_validation_reports_coco_metrics_with_the_gpu_matcher():
    """ref :219-250 end to end on the drop-in head with ``map_backend="gpu"``: the keys the reference logs, values equal
    to the oracle run on the head's own forward() output."""
    torch.manual_seed(0)
    CH, NCLS, SIZE, TOP = 16, 4, 128, 5
    head = ObjectDetection([3] + [CH] * TOP, NCLS, bottom_level=3, top_level=TOP, num_channels=CH, num_layers=1,
                           max_instances=20).to(DEV).eval()
    head.loc_head[-2].bias.data.fill_(0.0)
    head.map_backend = "gpu"
    g = torch.Generator().manual_seed(1)
    inputs = [torch.randn((3, 3, SIZE, SIZE), generator=g).to(DEV)] + [
        torch.randn((3, CH, SIZE // 2 ** l, SIZE // 2 ** l), generator=g).to(DEV) for l in range(1, TOP + 1)]
    gt = synth.gt_batch_np(3, 3, SIZE, SIZE, NCLS, 6, counts=[4, 0, 6])
    boxes = [_t(b) for b, _ in gt.per_image()]
    classes = [_t(c) for _, c in gt.per_image()]
    head.on_validation_start()
    assert isinstance(head.map_computer, metrics.DetectionMAP)
    with torch.no_grad():
        loss, _ = head.validation_step(inputs, classes, boxes)
        num, scores, pcls, pboxes = head.forward(inputs)
    out = head.on_validation_end()
    assert {"map", "map_50", "map_75", "map_small", "map_medium", "map_large", "mar_1", "mar_10", "mar_20", "mar_small",
            "mar_medium", "mar_large", "loss"} <= set(out)
    assert "classes" not in out and float(out["loss"]) == pytest.approx(float(loss), rel=1e-6)
    images = []
    for b in range(3):
        gb, gcl = gt.boxes[gt.offsets[b]:gt.offsets[b + 1]], gt.classes[gt.offsets[b]:gt.offsets[b + 1]]
        s, c, bx = scores[b].cpu().numpy(), pcls[b].cpu().numpy(), pboxes[b].cpu().numpy()
        w = mo.match_image(bx, s, c, gb, gcl)
        images.append(dict(scores=s[w["order"]], classes=c[w["order"]], dt_match=w["dt_match"], dt_ignore=w["dt_ignore"],
                           gt_classes=gcl, gt_ignore=w["gt_ignore"]))
    cats = sorted(set(int(v) for im in images for v in im["classes"]) | set(int(v) for im in images for v in im["gt_classes"]))
    want = mo.summarize(*mo.accumulate(images, cats, max_dets=(1, 10, 20)), max_dets=(1, 10, 20))
    for k, v in want.items():
        assert float(out[k]) == pytest.approx(v, abs=1e-6), k

"""oracle/torch_restatement.py against the REAL reference functions (container only: needs /root/reference).
This is the link that lets the gpu tests use the restatement as "the reference on the same device"."""
import pytest
import torch

from oracle import golden_cases as gc
from oracle import ref_loader
from oracle import torch_restatement as tr
from sihl_b200 import synth

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference tree only exists in the authoring container")


@pytest.mark.parametrize("name", sorted(gc.GEOMETRIES))
def test_offsets_scales_anchors(name):
    g = gc.GEOMETRIES[name]
    levels = gc.geom_levels(g)
    Ref = ref_loader.ObjectDetection()
    head = Ref(in_channels=[3] + [4] * g["top"], num_classes=1, bottom_level=g["bottom"], top_level=g["top"],
               num_channels=4, num_layers=0)
    inputs = [torch.zeros(1, 1, g["height"], g["width"])] + [None] * (g["bottom"] - 1) + [torch.zeros(1, 1, h, w) for h, w in levels]
    off, sc = head.get_offsets_and_scales(inputs)
    off_t, sc_t = tr.offsets_and_scales(levels, "cpu")
    assert torch.equal(off, off_t) and torch.equal(sc, sc_t)
    gold = gc.load("geom_" + name)
    assert torch.equal(tr.anchors_px(levels, g["width"], g["height"], "cpu"), torch.from_numpy(gold["anchors"]))


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("relative", [True, False])
def test_match_one_is_the_reference_bbox_matching(seed, relative):
    Ref = ref_loader.ObjectDetection()
    H, W = (640, 640) if seed % 2 == 0 else (384, 512)
    levels = synth.level_sizes(H, W)
    anchors = tr.anchors_px(levels, W, H, "cpu")
    gt = synth.gt_batch_np(1000 + seed, 3, H, W, 80, 60)
    for bx, _ in gt.per_image():
        t = torch.from_numpy(bx)
        ra, rv = Ref.bbox_matching(anchors, t, 9, relative=relative)
        a, v = tr.match_one(anchors, t, 9, relative)
        assert torch.equal(ra, a) and torch.equal(rv, v)


@pytest.mark.parametrize("name", ["train_cfg0", "train_test128", "train_cfg1", "train_empty"])
def test_train_losses_equal_the_reference_training_step(name):
    case = gc.TRAIN_CASES[name]
    g = gc.GEOMETRIES[case["geom"]]
    levels = gc.geom_levels(g)
    gold = gc.load(name)
    gt, maps = gc.case_gt(case), gc.train_maps(case)
    boxes = [torch.from_numpy(b) for b, _ in gt.per_image()]
    classes = [torch.from_numpy(c) for _, c in gt.per_image()]
    t = torch.from_numpy
    loss, metrics, _, _ = tr.train_losses(levels, g["width"], g["height"], boxes, classes, t(maps.loc_logits),
                                          t(maps.iou_preds), t(maps.box_raw), t(maps.cls_logits))
    for key in ("location_loss", "box_loss", "class_loss", "iou_loss"):
        want = float(gold[key])
        got = float(metrics[key])
        assert got == want or (got != got and want != want) or abs(got - want) <= 1e-6 * abs(want), (key, got, want)
    want = float(gold["loss"])
    assert float(loss) == want or abs(float(loss) - want) <= 1e-6 * abs(want) or (want != want)


@pytest.mark.parametrize("name", sorted(gc.FORWARD_CASES))
def test_forward_tail_equals_the_reference_forward(name):
    case = gc.FORWARD_CASES[name]
    g = gc.GEOMETRIES[case["geom"]]
    levels = gc.geom_levels(g)
    gold = gc.load(name)
    maps = gc.forward_maps(case)
    t = torch.from_numpy
    num, scores, cls, boxes, idx = tr.forward_tail(levels, g["width"], g["height"], t(maps.loc_logits), t(maps.box_raw),
                                                   t(maps.cls_logits), case["k"])
    assert torch.equal(num, t(gold["num_instances"])) and torch.equal(cls, t(gold["classes"]).long())
    assert torch.equal(scores, t(gold["scores"])) and torch.equal(boxes, t(gold["boxes"]))
    assert torch.equal(idx, t(gold["idx"]).long())


@pytest.mark.parametrize("name", sorted(gc.QUAD_CASES))
def test_quad_match_one_is_the_reference_quad_bbox_matching(name):
    """N1: tr.quad_match_one == QuadrilateralDetection.bbox_matching (ref quadrilateral_detection.py:266-294), raw
    outputs, and the sihl_b200 anchor helper == the reference's inline anchor construction (ref :156-161)."""
    Q = ref_loader.QuadrilateralDetection()
    case, gold = gc.QUAD_CASES[name], gc.load(name)
    anchors = torch.from_numpy(gold["anchors"])
    for bx, _ in gc.quad_gt(case).per_image():
        t = torch.from_numpy(bx).reshape(-1, 4)
        want = Q.bbox_matching(anchors, t, 9)
        got = tr.quad_match_one(anchors, t, 9)
        for w, g in zip(want, got):
            assert torch.equal(w, g)
    from sihl_b200.heads.quadrilateral_detection import quad_anchors
    sizes = [tuple(int(v) for v in s) for s in gold["level_sizes"]]
    mine = quad_anchors(sizes, range(case["bottom"], case["top"] + 1), case["top"], case["width"], case["height"], "cpu")
    assert torch.equal(mine, anchors)


def test_quads_to_boxes_equals_the_reference():
    """sihl_b200.heads.quadrilateral_detection.quads_to_boxes == QuadrilateralDetection.quads_to_boxes (ref :318-324)."""
    from sihl_b200.heads.quadrilateral_detection import quads_to_boxes
    Q = ref_loader.QuadrilateralDetection()
    g = torch.Generator().manual_seed(3)
    quads = torch.rand((17, 4, 2), generator=g) * 300
    assert torch.equal(quads_to_boxes(quads), Q.quads_to_boxes(quads))
    assert quads_to_boxes(quads[:0]).shape == (0, 4)

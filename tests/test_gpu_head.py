"""The drop-in head (sihl_b200.heads.ObjectDetection) on a GPU, mirroring the reference's own head tests
(ref: tests/heads/test_object_detection.py — shapes, loss >= 0, validation metrics) and adding what they never
check: values and gradients against the reference's operator sequence (oracle/torch_restatement.py) run on
the same device with the head's own MLP outputs."""
import numpy as np
import pytest
import torch

from oracle import torch_restatement as tr
from sihl_b200 import synth
from sihl_b200.heads import ObjectDetection

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
BATCH, CH, SIZE, BOTTOM, TOP, NCLS = 4, 32, 128, 3, 7, 16


def _model(**kw):
    torch.manual_seed(0)
    args = dict(in_channels=[3] + [CH] * TOP, num_classes=NCLS, bottom_level=BOTTOM, top_level=TOP, num_channels=CH, num_layers=2)
    args.update(kw)
    return ObjectDetection(**args).to(DEV)


def _backbone_output(batch=BATCH, size=SIZE):
    g = torch.Generator().manual_seed(1)
    return [torch.randn((batch, 3, size, size), generator=g).to(DEV)] + [
        torch.randn((batch, CH, size // 2 ** l, size // 2 ** l), generator=g).to(DEV) for l in range(1, TOP + 1)]


def _targets(batch=BATCH, size=SIZE):
    gt = synth.gt_batch_np(3, batch, size, size, NCLS, 6, counts=[0, 1, 4, 6][:batch])
    boxes = [torch.from_numpy(b).to(DEV) for b, _ in gt.per_image()]
    classes = [torch.from_numpy(c).to(DEV) for _, c in gt.per_image()]
    return {"classes": classes, "boxes": boxes}


def test_forward_shapes_like_the_reference_test():
    model = _model().eval()
    num_instances, scores, classes, boxes = model.forward(_backbone_output())
    assert tuple(num_instances.shape) == (BATCH,) and num_instances.dtype == torch.int64
    assert tuple(scores.shape) == (BATCH, model.max_instances)
    assert tuple(classes.shape) == (BATCH, model.max_instances) and classes.dtype == torch.int64
    assert tuple(boxes.shape) == (BATCH, model.max_instances, 4)
    assert (scores[:, :-1] >= scores[:, 1:]).all()


def test_training_and_validation_step_like_the_reference_test():
    model = _model()
    loss, metrics = model.training_step(_backbone_output(), **_targets())
    assert loss.item() >= 0 and set(metrics) == {"location_loss", "box_loss", "class_loss", "iou_loss"}
    model.on_validation_start()
    loss, _ = model.validation_step(_backbone_output(), **_targets())
    assert loss.item() >= 0
    out = model.on_validation_end()
    assert "loss" in out
    # the reference's own fixture: unsanitised integer boxes (x2 < x1 allowed), image 0 empty
    rng = np.random.RandomState(0)
    boxes = [torch.from_numpy(rng.randint(0, SIZE, (n, 4)).astype(np.float32)) for n in range(BATCH)]
    for b in boxes:
        b[(b[:, 0] == b[:, 2]) & (b[:, 1] == b[:, 3]), 2] += 3          # w == h == 0 is outside the defined domain
    classes = [torch.randint(0, NCLS, (n,)) for n in range(BATCH)]
    loss, _ = model.training_step(_backbone_output(), classes=[c.to(DEV) for c in classes], boxes=[b.to(DEV) for b in boxes])
    assert loss.item() >= 0 or torch.isinf(loss)                         # the reference test passes on +inf as well


def _dense_maps(model, inputs):
    flat = model._flat_feats(inputs)
    return (model.loc_head(flat).squeeze(2), model.iou_head(flat).squeeze(2), model.box_head(flat), model.cls_head(flat))


def test_training_step_values_and_gradients_match_the_reference_operators():
    model = _model().eval()                     # eval: BatchNorm laterals use running stats in both passes
    inputs, tgt = _backbone_output(), _targets()
    levels = model._level_sizes(inputs)
    loss, metrics = model.training_step(inputs, **tgt)
    loss.backward()
    ours = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
    model.zero_grad()
    loc, iou, box, cls = _dense_maps(model, inputs)
    ref_loss, ref_metrics, ref_assign, ref_rel = tr.train_losses(levels, SIZE, SIZE, tgt["boxes"], tgt["classes"], loc, iou, box, cls)
    ref_loss.backward()
    assert loss.item() == pytest.approx(ref_loss.item(), rel=2e-5)
    for k in metrics:
        assert metrics[k].item() == pytest.approx(ref_metrics[k].item(), rel=2e-5, abs=1e-7), k
    a, v = tr.canonical(ref_assign, ref_rel)
    assert torch.equal(model.last_assignment, a) and torch.equal(model.last_rel_iou, v)
    checked = 0
    for n, p in model.named_parameters():
        if p.grad is None:
            continue
        assert n in ours, n
        torch.testing.assert_close(ours[n], p.grad, rtol=2e-3, atol=2e-6, msg=lambda m, n=n: f"{n}: {m}")
        checked += 1
    assert checked > 20


def test_no_positive_early_out_matches_reference():
    model = _model().eval()
    inputs = _backbone_output(batch=2)
    empty = {"classes": [torch.zeros(0, dtype=torch.int64, device=DEV)] * 2, "boxes": [torch.zeros((0, 4), device=DEV)] * 2}
    loss, metrics = model.training_step(inputs, **empty)
    assert not torch.isfinite(loss)                                      # sum(bce) / 0 as in the reference (ref :163)
    assert metrics["box_loss"].item() == 0 and metrics["class_loss"].item() == 0 and metrics["iou_loss"].item() == 0


def test_forward_and_postprocess_match_the_reference_operators():
    model = _model().eval()
    inputs = _backbone_output()
    levels = model._level_sizes(inputs)
    with torch.no_grad():
        loc, iou, box, cls = _dense_maps(model, inputs)
        num, scores, classes, boxes = model.forward(inputs)
        t_num, t_scores, t_cls, t_boxes, _ = tr.forward_tail(levels, SIZE, SIZE, loc, box, cls, model.max_instances)
        assert torch.equal(num, t_num)
        torch.testing.assert_close(scores, t_scores, rtol=1e-5, atol=1e-7)
        assert (classes == t_cls).float().mean() > 0.99                  # gathered-row vs dense GEMM: last-bit logit noise
        torch.testing.assert_close(boxes, t_boxes, rtol=1e-4, atol=1e-3)
        # extension: dense decode + class-aware NMS; bias the location head so that candidates exist
        model.loc_head[-2].bias.data.fill_(-1.0)
        loc, iou, box, cls = _dense_maps(model, inputs)
        p_num, p_scores, p_cls, p_boxes = model.postprocess(inputs, 0.05, 0.5)
        r_num, r_scores, r_cls, r_boxes = tr.dense_postprocess(levels, SIZE, SIZE, loc, box, cls, 0.05, 0.5, model.max_instances)
        assert torch.equal(p_num, r_num) and p_num.max() > 0
        torch.testing.assert_close(p_scores, r_scores, rtol=1e-5, atol=1e-7)
        assert torch.equal(p_cls, r_cls)
        torch.testing.assert_close(p_boxes, r_boxes, rtol=1e-4, atol=1e-3)


def test_autocast_bf16_and_static_bbox_matching():
    model = _model()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss, _ = model.training_step(_backbone_output(), **_targets())
    assert torch.isfinite(loss)
    loss.backward()
    assert all(torch.isfinite(p.grad).all() for p in model.parameters() if p.grad is not None)
    levels = synth.level_sizes(SIZE, SIZE, mode="floor")
    anchors = tr.anchors_px(levels, SIZE, SIZE, DEV)
    gt = _targets()["boxes"][3]
    a, v = ObjectDetection.bbox_matching(anchors, gt, 9, relative=True)
    ra, rv = tr.canonical(*tr.match_one(anchors, gt, 9, True))
    assert torch.equal(a, ra) and torch.equal(v, rv)


def test_get_saliency_and_offsets():
    model = _model().eval()
    inputs = _backbone_output()
    sal = model.get_saliency(inputs)
    assert tuple(sal.shape) == (BATCH, SIZE // 8, SIZE // 8) and (sal >= 0).all() and (sal <= 1).all()
    off, sc = model.get_offsets_and_scales(inputs)
    t_off, t_sc = tr.offsets_and_scales(model._level_sizes(inputs), DEV)
    assert torch.equal(off, t_off) and torch.equal(sc, t_sc)


def test_training_assign_with_more_images_than_fit_the_kernel_parameters():
    """Above SIHL_OD_MAX_BATCH_BY_VALUE (512) images the per-image gt counts cannot ride in the kernel parameters: the
    offsets are uploaded instead.  Same assignment as two half-size calls; the device-offsets form (CUDA-graph replayable)
    gives the same result as the host-counts form."""
    from sihl_b200 import ops
    size, B, C = 64, 600, 3
    levels = synth.level_sizes(size, size, 3, 5)
    counts = [(i * 7) % 4 for i in range(B)]                      # 0..3 gts per image, some images empty
    gt = synth.gt_batch_np(11, B, size, size, C, 3, counts=counts)
    gb, gc_ = torch.from_numpy(gt.boxes).to(DEV), torch.from_numpy(gt.classes).to(DEV)
    st = ops.train_assign(levels, size, size, gb, gc_, counts, B, 9)                         # B > 512: upload path
    half = B // 2
    cut = int(gt.offsets[half])
    a = ops.train_assign(levels, size, size, gb[:cut], gc_[:cut], counts[:half], half, 9)    # by-value path
    b = ops.train_assign(levels, size, size, gb[cut:], gc_[cut:], counts[half:], half, 9)
    assert torch.equal(st.assignment, torch.cat([a.assignment, b.assignment]))
    assert torch.equal(st.rel_iou, torch.cat([a.rel_iou, b.rel_iou]))
    assert int(st.pos_total) == int(a.pos_total) + int(b.pos_total) > 0
    dev_form = ops.train_assign(levels, size, size, gb[:cut], gc_[:cut], None, half, 9,
                                gt_offsets=torch.from_numpy(gt.offsets[:half + 1].copy()).to(DEV))
    assert torch.equal(dev_form.assignment, a.assignment) and torch.equal(dev_form.pos_index, a.pos_index)
    P = int(a.pos_total)
    assert (a.pos_index[P:] == 0).all() and (a.pos_index[1:P] > a.pos_index[:P - 1]).all()   # ascending rows, zero padding

"""X1 — the drop-in head under the reference's OWN caller, with the REAL reference as the same-device oracle.

BASELINE.json configs[0] as written: resnet18 torchvision backbone + FPN, 320x320, batch 2, 10 classes, <= 20 gt per
image.  ``SihlModel(TorchvisionBackbone("resnet18"), FPN, [head])`` is built twice from the reference's unmodified source
(``oracle/ref_loader.py``: /root/reference in the authoring container, the staged copy ``oracle/_ref/sihl_src`` on the
GPU box) — once around the reference ``ObjectDetection`` and once around ``sihl_b200.heads.ObjectDetection`` loaded with
the same ``state_dict`` — and driven exactly as ``SihlLightningModule`` drives it (ref lightning_module.py:88,95-96,
145-146; sihl_model.py:17-25): ``extract_features`` -> ``head.training_step(inputs, **target)`` / ``validation_step`` /
``forward``; ``get_saliency`` as ``sihl.visualization`` calls it (ref visualization/object_detection.py:25).
"""
import copy

import numpy as np
import pytest
import torch

from oracle import ref_loader
from sihl_b200 import synth
from sihl_b200.heads import ObjectDetection

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_loader.available(),
                                 reason="reference source neither at /root/reference nor staged in oracle/_ref/sihl_src "
                                        "(python -m oracle.stage_reference)")]
DEV = "cuda:0"
SIZE, BATCH, NCLS, BOTTOM, TOP = 320, 2, 10, 3, 7
COUNTS = [20, 7]


def _build(seed=0):
    """(reference model, our model): shared backbone + neck weights, identical head weights."""
    torch.manual_seed(seed)
    Model, Backbone, Neck, RefHead = ref_loader.SihlModel(), ref_loader.TorchvisionBackbone(), ref_loader.FPN(), \
        ref_loader.ObjectDetection()
    backbone = Backbone("resnet18", top_level=5)
    neck = Neck(backbone.out_channels, 256, bottom_level=BOTTOM, top_level=TOP)
    ref_head = RefHead(neck.out_channels, NCLS, bottom_level=BOTTOM, top_level=TOP)
    # a trained location head fires somewhere: move the bias so that forward() sees scores on both sides of 0.5
    ref_head.loc_head[-2].bias.data.fill_(-0.2)
    ref = Model(backbone, neck, [ref_head]).to(DEV)
    head = ObjectDetection(neck.out_channels, NCLS, bottom_level=BOTTOM, top_level=TOP)
    missing = head.load_state_dict(ref_head.state_dict(), strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    ours = Model(copy.deepcopy(backbone), copy.deepcopy(neck), [head]).to(DEV)
    ours.load_state_dict(ref.state_dict())
    return ref, ours


def _batch(seed=5):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn((BATCH, 3, SIZE, SIZE), generator=g).to(DEV)
    gt = synth.gt_batch_np(seed, BATCH, SIZE, SIZE, NCLS, max(COUNTS), counts=COUNTS)
    target = {"classes": [torch.from_numpy(c).to(DEV) for _, c in gt.per_image()],
              "boxes": [torch.from_numpy(b).to(DEV) for b, _ in gt.per_image()]}
    return x, target


def _train_step(model, x, target, autocast=None):
    """What SihlLightningModule.training_step does with one head (ref lightning_module.py:88-98)."""
    model.zero_grad(set_to_none=True)
    ctx = torch.autocast("cuda", dtype=autocast) if autocast is not None else torch.autocast("cuda", enabled=False)
    with ctx:
        head_inputs = model.extract_features(x)
        loss, metrics = model.heads[0].training_step(head_inputs, **target)
    loss.backward()
    grads = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
    return loss.detach(), {k: v.detach() for k, v in metrics.items()}, grads, head_inputs


def test_config0_training_step_matches_the_reference_head():
    ref, ours = _build()
    ref.train(), ours.train()
    x, target = _batch()
    r_loss, r_metrics, r_grads, r_inputs = _train_step(ref, x, target)
    o_loss, o_metrics, o_grads, o_inputs = _train_step(ours, x, target)
    assert torch.isfinite(r_loss)
    assert o_loss.item() == pytest.approx(r_loss.item(), rel=1e-5)
    assert set(o_metrics) == set(r_metrics)
    for k in r_metrics:
        assert o_metrics[k].item() == pytest.approx(r_metrics[k].item(), rel=1e-5, abs=1e-7), k
    # assignment and relative IoU: bit-equal to the REAL bbox_matching on the same device (ref :143-148, :252-284)
    head, ref_head = ours.heads[0], ref.heads[0]
    offsets, scales = ref_head.get_offsets_and_scales(r_inputs)
    anchors = (offsets + scales) * torch.tensor([[SIZE, SIZE, SIZE, SIZE]], device=DEV)
    for b in range(BATCH):
        a, v = ref_head.bbox_matching(anchors, target["boxes"][b], ref_head.topk, relative=True)
        assert torch.equal(head.last_rel_iou[b], v), f"rel_iou differs in image {b}"
        pos = v > 0
        assert torch.equal(head.last_assignment[b][pos], a[pos]), f"assignment differs in image {b}"
        assert (head.last_assignment[b][~pos] == -1).all()
        assert int(pos.sum()) > 0
    # parameter gradients of backbone, neck and head (fp32: GEMMs on a different number of gathered rows may pick
    # another cuBLAS kernel, hence a tolerance instead of bit equality)
    assert set(o_grads) == set(r_grads)
    worst = 0.0
    for n, g in r_grads.items():
        scale = g.abs().max().item()
        if scale == 0:
            assert o_grads[n].abs().max().item() < 1e-12, n
            continue
        err = (o_grads[n] - g).abs().max().item() / scale
        worst = max(worst, err)
        assert err < 5e-3, f"{n}: {err}"        # bn1.bias of the backbone (batch of 2, 5 BN-normalised scales below it): 2.1e-3
    assert len(r_grads) > 100, len(r_grads)


def test_config0_forward_and_saliency_match_the_reference_head():
    ref, ours = _build()
    ref.eval(), ours.eval()
    x, _ = _batch()
    with torch.no_grad():
        r_num, r_scores, r_cls, r_boxes = ref(x)[0]                        # SihlModel.forward, ref sihl_model.py:23-25
        o_num, o_scores, o_cls, o_boxes = ours(x)[0]
        assert o_num.dtype == r_num.dtype and o_cls.dtype == r_cls.dtype
        assert torch.equal(o_num, r_num)
        torch.testing.assert_close(o_scores, r_scores, rtol=1e-6, atol=1e-7)
        assert torch.equal(o_cls, r_cls)
        torch.testing.assert_close(o_boxes, r_boxes, rtol=1e-5, atol=1e-3)
        feats = ref.extract_features(x)
        r_sal = ref.heads[0].get_saliency(feats)                            # ref :70-81, values (not only the range)
        o_sal = ours.heads[0].get_saliency(ours.extract_features(x))
        assert o_sal.shape == r_sal.shape
        torch.testing.assert_close(o_sal, r_sal, rtol=1e-6, atol=1e-7)
        r_off, r_sc = ref.heads[0].get_offsets_and_scales(feats)
        o_off, o_sc = ours.heads[0].get_offsets_and_scales(feats)
        assert torch.equal(o_off, r_off) and torch.equal(o_sc, r_sc)


def test_config0_validation_step_feeds_the_metric_like_the_reference():
    """ref :227-250 with the recording stand-in for torchmetrics' MeanAveragePrecision (torchmetrics is not installed):
    both heads must hand the same detections and targets to ``update`` and report the same running loss."""
    ref, ours = _build()
    ref.eval(), ours.eval()
    x, target = _batch()
    outs = []
    for model in (ref, ours):
        head = model.heads[0]
        head.on_validation_start()
        with torch.no_grad():
            inputs = model.extract_features(x)                             # ref lightning_module.py:138,145-146
            loss, metrics = head.validation_step(inputs, **target)
            loss2, _ = head.validation_step(inputs, **target)
        end = head.on_validation_end()
        outs.append((loss, metrics, head.map_computer, end))
    (r_loss, r_metrics, r_map, r_end), (o_loss, o_metrics, o_map, o_end) = outs
    assert o_loss.item() == pytest.approx(r_loss.item(), rel=1e-5)
    assert set(o_end) == set(r_end) and "loss" in o_end
    assert float(o_end["loss"]) == pytest.approx(float(r_end["loss"]), rel=1e-5)
    assert len(o_map.preds) == len(r_map.preds) == 2 * BATCH and len(o_map.targets) == len(r_map.targets)
    for po, pr in zip(o_map.preds, r_map.preds):
        assert set(po) == set(pr) == {"scores", "labels", "boxes"}
        torch.testing.assert_close(po["scores"], pr["scores"], rtol=1e-6, atol=1e-7)
        assert torch.equal(po["labels"], pr["labels"])
        torch.testing.assert_close(po["boxes"], pr["boxes"], rtol=1e-5, atol=1e-3)
    for to, tr_ in zip(o_map.targets, r_map.targets):
        assert torch.equal(to["labels"], tr_["labels"]) and torch.equal(to["boxes"], tr_["boxes"])


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_config0_training_step_under_autocast_matches_the_reference_head(dtype):
    """precision="16-mixed" (ref examples/object_detection.py:294).  The reference evaluates the location BCE on the
    HALF logits (ref :160-161: no .to(float32)), the other three losses after an upcast; the kernels reproduce both
    (half maps are loaded as they are and upcast in registers; the log-sigmoid term is rounded to the map type)."""
    ref, ours = _build()
    ref.train(), ours.train()
    x, target = _batch()
    r_loss, r_metrics, r_grads, _ = _train_step(ref, x, target, autocast=dtype)
    o_loss, o_metrics, o_grads, _ = _train_step(ours, x, target, autocast=dtype)
    assert torch.isfinite(r_loss) and torch.isfinite(o_loss)
    # loc_head / iou_head see identical inputs and shapes in both models -> identical half logits: tight tolerance
    assert o_metrics["location_loss"].item() == pytest.approx(r_metrics["location_loss"].item(), rel=2e-5)
    assert o_metrics["iou_loss"].item() == pytest.approx(r_metrics["iou_loss"].item(), rel=2e-5)
    # box / class heads run on a different number of rows (static capacity) -> half-precision GEMM noise
    eps = 2e-2 if dtype == torch.bfloat16 else 4e-3
    assert o_metrics["box_loss"].item() == pytest.approx(r_metrics["box_loss"].item(), rel=eps)
    assert o_metrics["class_loss"].item() == pytest.approx(r_metrics["class_loss"].item(), rel=eps)
    assert o_loss.item() == pytest.approx(r_loss.item(), rel=eps)
    assert set(o_grads) == set(r_grads)
    for n in ("heads.0.loc_head.0.weight", "heads.0.iou_head.0.weight", "heads.0.cls_head.0.weight",
              "heads.0.box_head.0.weight", "backbone.model.conv1.weight"):
        g = r_grads[n]
        err = (o_grads[n] - g).abs().max().item() / max(g.abs().max().item(), 1e-30)
        assert err < (0.15 if dtype == torch.bfloat16 else 0.05), f"{n}: {err}"


def test_empty_batch_early_out_matches_the_reference_head():
    """ref :165-172 — no ground truth at all: the loss is the location loss alone (sum(bce) / 0 = inf in both)."""
    ref, ours = _build()
    ref.eval(), ours.eval()
    x, _ = _batch()
    empty = {"classes": [torch.zeros(0, dtype=torch.int64, device=DEV)] * BATCH,
             "boxes": [torch.zeros((0, 4), device=DEV)] * BATCH}
    with torch.no_grad():
        r_loss, r_metrics = ref.heads[0].training_step(ref.extract_features(x), **empty)
        o_loss, o_metrics = ours.heads[0].training_step(ours.extract_features(x), **empty)
    assert torch.isinf(r_loss) and torch.isinf(o_loss)
    for k in ("box_loss", "class_loss", "iou_loss"):
        assert float(o_metrics[k]) == float(r_metrics[k]) == 0.0


def test_config0_a_few_optimizer_steps_follow_the_reference_trajectory():
    """What the Lightning loop does without Lightning (it is not installed): four SGD steps on both models from the same
    initialisation and the same batches (ref lightning_module.py:68-119 -> optimizer step).  The losses must stay
    together step after step — gradients flow through the replacement exactly as through the reference head."""
    ref, ours = _build()
    ref.train(), ours.train()
    opt_r = torch.optim.SGD(ref.parameters(), lr=2e-3, momentum=0.9)
    opt_o = torch.optim.SGD(ours.parameters(), lr=2e-3, momentum=0.9)
    trace = []
    for step in range(4):
        x, target = _batch(seed=20 + step)
        losses = []
        for model, opt in ((ref, opt_r), (ours, opt_o)):
            opt.zero_grad(set_to_none=True)
            loss, _ = model.heads[0].training_step(model.extract_features(x), **target)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(model.parameters(), 0.1)      # ref examples/object_detection.py:293 (grad-clip 0.1)
            opt.step()
            losses.append(loss.item())
        trace.append(losses)
        assert losses[1] == pytest.approx(losses[0], rel=5e-4), (step, trace)
    assert trace[-1][0] != trace[0][0]                                    # the models did move

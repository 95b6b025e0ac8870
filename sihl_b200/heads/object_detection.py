"""``ObjectDetection`` head — drop-in for ``sihl.heads.ObjectDetection`` with the dense,
non-learned tail running in ``libsihl_b200.so``.

Mirrors the reference class (``/root/reference/src/sihl/heads/object_detection.py``):
same constructor (:14-23), attributes, ``state_dict`` keys (``laterals.*``, ``loc_head.*``,
``cls_head.*``, ``box_head.*``, ``iou_head.*``), the ``Head`` protocol methods
(``src/sihl/heads/__init__.py:28-53``) with the same signatures, return types and error
behaviour, so ``SihlLightningModule`` (``src/sihl/lightning_module.py:95-98,145-148``),
``SihlModel`` and ``sihl.visualization`` keep working unchanged.

What stays PyTorch: the learned part (1x1 conv + BN laterals, four LayerNorm/SiLU MLPs —
dense contractions, out of scope here).  What moved to CUDA kernels behind
the C ABI: anchor tables (ref :83-97,:134-140), the per-image ``bbox_matching`` loop
(:143-148,:252-284), the four loss reductions and their backward (:157-210), the top-k decode
of ``forward`` (:108-121), plus the north-star extension ``postprocess`` (dense decode +
class-aware NMS; the reference has no NMS).

Differences a caller can observe, all documented in DESIGN.md:
  * tensors must live on a CUDA device (there is no CPU path);
  * ``assignment`` is canonical: -1 wherever ``rel_iou`` is not > 0;
  * exact ties inside ``topk`` resolve to the lowest index (torch leaves them undefined);
  * ONNX export of ``forward`` is not available (custom kernels).
"""
from __future__ import annotations

from functools import partial
from typing import Dict, List, Optional, Tuple

import torch
from torch import Tensor, nn
from torch.nn import functional
from torchvision import ops as tvops

from .. import ops

_TOTAL_WEIGHTS = (1.0, 10.0, 1.0, 1.0)          # loss = loc + 10*box + class + iou, ref :210


class _DetectionLoss(torch.autograd.Function):
    """[location, box, class, iou, total] losses of ref :157-210 from the head outputs.

    forward: k_dense_loss + k_pos_loss + k_loss_finalize; backward: k_dense_loss_bwd +
    k_pos_loss_bwd.  ``assignment`` / ``rel_iou`` carry no gradient (they are
    functions of the anchors and the ground truth only)."""

    @staticmethod
    def forward(ctx, loc_logits, iou_preds, box_raw, cls_logits, state):
        s = state
        sums = ops.new_sums(loc_logits.device)
        loc32 = loc_logits.detach().float().contiguous()
        iou32 = None if iou_preds is None else iou_preds.detach().float().contiguous()
        ops.dense_loss(loc32, iou32, s["rel_iou"], sums)
        box32 = cls32 = None
        if s["P"] > 0:
            box32 = box_raw.detach().float().contiguous()
            cls32 = cls_logits.detach().float().contiguous()
            ops.pos_loss(s["pos_index"], None, s["P"], s["A"], s["rel_iou"], s["assignment"], s["offsets"], s["scales"],
                         s["img_w"], s["img_h"], s["gt"], box32, cls32, False, sums)
        if s.get("reduce_sums") is not None:
            s["reduce_sums"](sums)                      # global-batch normalisers across ranks (dist.py)
        ctx.state, ctx.sums = s, sums
        ctx.saved = (loc32, iou32, box32, cls32)
        ctx.in_dtypes = tuple(None if t is None else t.dtype for t in (loc_logits, iou_preds, box_raw, cls_logits))
        return ops.loss_finalize(sums)

    @staticmethod
    def backward(ctx, grad_out):
        s, sums = ctx.state, ctx.sums
        loc32, iou32, box32, cls32 = ctx.saved
        w = torch.tensor(_TOTAL_WEIGHTS, dtype=torch.float32, device=grad_out.device)
        grad_terms = (grad_out[:4].float() + grad_out[4].float() * w).contiguous()
        dloc, diou = ops.dense_loss_bwd(loc32, iou32, s["rel_iou"], sums, grad_terms,
                                        want_dloc=ctx.needs_input_grad[0], want_diou=ctx.needs_input_grad[1])
        dbox = dcls = None
        if s["P"] > 0 and (ctx.needs_input_grad[2] or ctx.needs_input_grad[3]):
            dbox, dcls = ops.pos_loss_bwd(s["pos_index"], None, s["P"], s["A"], s["rel_iou"], s["assignment"], s["offsets"],
                                          s["scales"], s["img_w"], s["img_h"], s["gt"], box32, cls32, False, sums, grad_terms)
        cast = lambda g, dt: None if g is None or dt is None else g.to(dt)
        return (cast(dloc, ctx.in_dtypes[0]), cast(diou, ctx.in_dtypes[1]), cast(dbox, ctx.in_dtypes[2]),
                cast(dcls, ctx.in_dtypes[3]), None)


class ObjectDetection(nn.Module):
    def __init__(
        self,
        in_channels: List[int],
        num_classes: int,
        bottom_level: int = 3,
        top_level: int = 5,
        num_channels: int = 256,
        num_layers: int = 4,
        max_instances: int = 100,
    ) -> None:
        """Same arguments and checks as the reference constructor (ref :14-39)."""
        assert num_classes > 0, num_classes
        assert len(in_channels) > top_level, (len(in_channels), top_level)
        assert 0 < bottom_level <= top_level, (bottom_level, top_level)
        assert num_channels % 4 == 0, num_channels
        assert num_layers >= 0, num_layers
        assert max_instances > 0, max_instances
        super().__init__()

        self.in_channels = in_channels
        self.num_classes = num_classes
        self.bottom_level, self.top_level = bottom_level, top_level
        self.levels = range(bottom_level, top_level + 1)
        self.num_channels = num_channels
        self.num_layers = num_layers
        self.max_instances = max_instances
        self.topk = 9

        # learned part: identical modules => identical state_dict keys and initialisation (ref :51-61)
        mlp = partial(tvops.MLP, norm_layer=nn.LayerNorm, activation_layer=nn.SiLU)
        conv = partial(tvops.Conv2dNormActivation, activation_layer=None)
        self.laterals = nn.ModuleList([conv(in_channels[level], num_channels, 1) for level in self.levels])
        hidden = [num_channels] * num_layers
        self.loc_head = mlp(num_channels, hidden + [1])
        self.loc_head[-2].bias.data.fill_(-5.0)
        self.cls_head = mlp(num_channels, hidden + [num_classes])
        self.box_head = mlp(num_channels, hidden + [4])
        self.iou_head = mlp(num_channels, hidden + [1])

        self.output_shapes = {
            "num_instances": ("batch_size",),
            "scores": ("batch_size", max_instances),
            "classes": ("batch_size", max_instances),
            "boxes": ("batch_size", max_instances, 4),
        }
        # "local": normalise by this process's batch (what the reference does, also under DDP);
        # "global": all-reduce the 8 partial sums first (equals one process over the global batch).
        self.loss_reduction = "local"
        self.process_group = None

    # ------------------------------------------------------------------ helpers
    def _level_sizes(self, inputs: List[Tensor]) -> List[Tuple[int, int]]:
        return [tuple(int(v) for v in inputs[level].shape[2:]) for level in self.levels]

    def _flat_feats(self, inputs: List[Tensor]) -> Tensor:
        feats = [lateral(inputs[level]) for level, lateral in zip(self.levels, self.laterals)]     # ref :102-105
        return torch.cat([x.flatten(2).transpose(1, 2) for x in feats], 1)

    def get_offsets_and_scales(self, inputs: List[Tensor]) -> Tuple[Tensor, Tensor]:
        """ref :83-97 — one cached kernel launch instead of ~40 ATen launches per call."""
        device = inputs[0].device
        height, width = inputs[0].shape[2:]
        offsets, scales, _ = ops.anchor_tables(self._level_sizes(inputs), int(width), int(height), device)
        return offsets, scales

    def get_saliency(self, inputs: List[Tensor]) -> Tensor:
        """ref :70-81 (visualisation only; learned part, stays PyTorch)."""
        device = inputs[self.bottom_level].device
        batch_size, _, full_height, full_width = inputs[self.bottom_level].shape
        output = torch.zeros((batch_size, full_height, full_width), device=device)
        for lateral, level in zip(self.laterals, self.levels):
            height, width = inputs[level].shape[2:]
            feats = lateral(inputs[level]).flatten(2).transpose(1, 2)
            scores = self.loc_head(feats).sigmoid().transpose(1, 2).reshape(batch_size, 1, height, width)
            scores = functional.interpolate(scores, size=(full_height, full_width))
            output = torch.maximum(output, scores.squeeze(1))
        return output

    # ------------------------------------------------------------------ inference
    def forward(self, inputs: List[Tensor]) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
        """ref :99-122 -> (num_instances i64 [B], scores [B,K], classes i64 [B,K], boxes [B,K,4] px)."""
        (batch_size, _, height, width), device = inputs[0].shape, inputs[0].device
        flat_feats = self._flat_feats(inputs)
        offsets, scales, _ = ops.anchor_tables(self._level_sizes(inputs), int(width), int(height), device)
        loc_logits = self.loc_head(flat_feats).squeeze(2)                                   # ref :108
        top_logits, loc_idxs = ops.topk_locations(loc_logits.detach().float(), self.max_instances)   # ref :109
        rows = torch.arange(batch_size, device=device).view(batch_size, 1)
        top_feats = flat_feats[rows, loc_idxs]                                              # ref :112
        class_logits = self.cls_head(top_feats)                                             # ref :116
        box_raw = self.box_head(top_feats)                                                  # ref :121
        return ops.decode_rows(top_logits, loc_idxs, class_logits.detach().float(), box_raw.detach().float(),
                               offsets, scales, int(width), int(height))

    @torch.no_grad()
    def postprocess(self, inputs: List[Tensor], score_threshold: float = 0.05, iou_threshold: float = 0.5
                    ) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
        """North-star extension (the reference has no NMS): every location is decoded with the
        score/class semantics of ``forward`` (score = sigmoid(location logit), class = argmax),
        thresholded, and de-duplicated by class-aware NMS; same output format as ``forward``."""
        (_, _, height, width) = inputs[0].shape
        flat_feats = self._flat_feats(inputs)
        loc_logits = self.loc_head(flat_feats).squeeze(2).float()
        cls_logits = self.cls_head(flat_feats).float()
        box_raw = self.box_head(flat_feats).float()
        return ops.dense_postprocess(loc_logits, cls_logits, box_raw, self._level_sizes(inputs), int(width), int(height),
                                     score_threshold, iou_threshold, self.max_instances)

    # ------------------------------------------------------------------ training
    def _reduce_sums(self) -> Optional[callable]:
        if self.loss_reduction != "global":
            return None
        from .. import dist
        return partial(dist.all_reduce_sums, group=self.process_group)

    def training_step(
        self,
        inputs: List[Tensor],
        classes: List[Tensor],
        boxes: List[Tensor],
        is_validating: bool = False,
    ) -> Tuple[Tensor, Dict[str, float]]:
        """ref :124-217 -> (loss, {"location_loss", "box_loss", "class_loss", "iou_loss"})."""
        assert len(inputs) > self.top_level, "too few input levels"
        device = inputs[0].device
        batch_size, _, full_height, full_width = inputs[0].shape
        levels = self._level_sizes(inputs)
        width, height = int(full_width), int(full_height)

        # anchors + assignment: functions of shapes and gt only (ref :139-148) -> before the MLPs
        offsets, scales, anchors = ops.anchor_tables(levels, width, height, device)
        num_anchors = anchors.shape[0]
        gt = ops.GtBatch.from_lists(boxes, classes, device)
        assert gt.batch_size == batch_size, (gt.batch_size, batch_size)
        sel = ops.assign_select(anchors, levels, width, height, gt, self.topk,
                                terms=ops.anchor_terms(levels, width, height, device))
        res = ops.assign_resolve(sel, gt, num_anchors, self.topk, True, want_positives=True)
        pos_index, pos_total, _ = ops.pos_compact(res["tile_pos_count"], res["tile_pos_rows"], batch_size, num_anchors)
        # the one host sync of the step: P sizes the gathered rows (the reference syncs ~7x per image)
        num_pos = int(pos_total.item())
        pos_index = pos_index[:num_pos]

        flat_feats = self._flat_feats(inputs)                                               # ref :151-154
        loc_logits = self.loc_head(flat_feats).squeeze(2)                                   # ref :157
        state = dict(rel_iou=res["iou"], assignment=res["assignment"], pos_index=pos_index, P=num_pos, A=num_anchors,
                     offsets=offsets, scales=scales, img_w=width, img_h=height, gt=gt, reduce_sums=self._reduce_sums())
        if num_pos == 0:                                                                    # ref :165-172
            out = _DetectionLoss.apply(loc_logits, None, None, None, state)
        else:
            iou_preds = self.iou_head(flat_feats).squeeze(2)                                # ref :175
            o2m_feats = flat_feats.reshape(batch_size * num_anchors, -1)[pos_index.long()]  # ref :184
            box_raw = self.box_head(o2m_feats)                                              # ref :189
            class_logits = self.cls_head(o2m_feats)                                         # ref :200
            out = _DetectionLoss.apply(loc_logits, iou_preds, box_raw, class_logits, state)
        self.last_assignment, self.last_rel_iou = res["assignment"], res["iou"]
        metrics = {"location_loss": out[0], "box_loss": out[1], "class_loss": out[2], "iou_loss": out[3]}
        return out[4], metrics

    training_loss = training_step          # north-star name for the same entry point

    # ------------------------------------------------------------------ validation
    def on_validation_start(self) -> None:
        """ref :219-225.  torchmetrics is optional here: without it the mAP is skipped and the
        loss is averaged by a two-scalar running mean."""
        try:
            from torchmetrics import MeanMetric
            from torchmetrics.detection.mean_ap import MeanAveragePrecision
            self.loss_computer = MeanMetric(nan_strategy="ignore")
            thresholds = [1, min(self.max_instances, 10), self.max_instances]
            self.map_computer = MeanAveragePrecision(max_detection_thresholds=thresholds, backend="faster_coco_eval")
        except Exception:
            self.loss_computer = _RunningMean()
            if hasattr(self, "map_computer"):
                del self.map_computer

    def validation_step(self, inputs: List[Tensor], classes: List[Tensor], boxes: List[Tensor]
                        ) -> Tuple[Tensor, Dict[str, float]]:
        """ref :227-240."""
        num_instances, scores, pred_classes, pred_boxes = self.forward(inputs)
        if hasattr(self, "map_computer"):
            self.map_computer.to(scores.device).update(
                [{"scores": s, "labels": c, "boxes": b} for s, c, b in zip(scores, pred_classes, pred_boxes)],
                [{"labels": c, "boxes": b} for c, b in zip(classes, boxes)],
            )
        loss, metrics = self.training_step(inputs, classes, boxes, is_validating=True)
        self.loss_computer.to(loss.device).update(loss)
        return loss, metrics

    def on_validation_end(self) -> Dict[str, float]:
        """ref :242-250."""
        metrics = {}
        if hasattr(self, "map_computer"):
            metrics = self.map_computer.compute()
            for key in list(metrics.keys()):
                if "per_class" in key or key in {"classes", "ious"}:
                    del metrics[key]
        metrics["loss"] = self.loss_computer.compute()
        return metrics

    # ------------------------------------------------------------------ assigner (also used by sibling heads)
    @staticmethod
    def bbox_matching(anchors: Tensor, gt_boxes: Tensor, topk: int, relative: bool = False) -> Tuple[Tensor, Tensor]:
        """ref :252-284 — ``(assignment i64 [A], iou f32 [A])`` for one image, arbitrary anchors
        (``InstanceSegmentation`` / ``KeypointDetection`` call this verbatim)."""
        return ops.bbox_matching(anchors, gt_boxes, topk, relative)


class _RunningMean:
    """Stand-in for ``torchmetrics.MeanMetric(nan_strategy="ignore")`` when torchmetrics is absent."""

    def __init__(self) -> None:
        self.total, self.count = 0.0, 0

    def to(self, device):
        return self

    def update(self, value: Tensor) -> None:
        v = float(value.detach()) if hasattr(value, "detach") else float(value)
        if v == v:                       # ignore NaN
            self.total, self.count = self.total + v, self.count + 1

    def compute(self) -> Tensor:
        return torch.tensor(self.total / self.count if self.count else float("nan"))

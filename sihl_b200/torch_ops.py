"""``torch.library`` registration of the detection-head ops — the "torch custom-op shim only for tensor handoff" of
BASELINE.json's north_star / SURVEY.md §7.

Every op is a thin wrapper around the ctypes hand-off in :mod:`sihl_b200.ops` (which in turn only passes raw device
pointers to ``libsihl_b200.so``): no arithmetic here.  What the registration adds:

* **fake (meta) implementations** — output shapes / dtypes without touching the GPU, so the head traces under
  ``FakeTensorMode``, ``torch.export`` and ``torch.compile(fullgraph=True)`` (the reference's ``forward`` is exported
  through dynamo in its own test, ref tests/heads/test_object_detection.py:83-107);
* **autograd** for the loss op (``register_autograd``): backward is the ``train_loss_bwd`` op, SURVEY.md §7.4.

The eager head calls :mod:`sihl_b200.ops` directly (a ``custom_op`` dispatch costs tens of microseconds, which matters
for a 100 µs training tail); under ``torch.compiler.is_compiling()`` it routes through these ops instead — same
kernels, same results.

Level sizes travel as a flat ``int[]`` ``[h0, w0, h1, w1, ...]`` (schema types have no list of pairs).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
from torch import Tensor

from . import ops

_F32, _I64, _I32 = torch.float32, torch.int64, torch.int32


def flat_levels(levels) -> List[int]:
    return [int(v) for hw in levels for v in hw]


def _pairs(level_hw: List[int]):
    return [(int(level_hw[i]), int(level_hw[i + 1])) for i in range(0, len(level_hw), 2)]


def _num_anchors(level_hw: List[int]) -> int:
    return sum(h * w for h, w in _pairs(level_hw))


# --------------------------------------------------------------------------- a1/a2
@torch.library.custom_op("sihl_b200::anchor_tables", mutates_args=())
def anchor_tables(like: Tensor, level_hw: List[int], img_w: int, img_h: int) -> Tuple[Tensor, Tensor, Tensor]:
    """ref :83-97, :134-140 -> (offsets, scales, anchors) [A,4] on ``like``'s device (copies of the cached tables)."""
    o, s, a = ops.anchor_tables(_pairs(level_hw), img_w, img_h, like.device)
    return o.clone(), s.clone(), a.clone()


@anchor_tables.register_fake
def _(like, level_hw, img_w, img_h):
    A = _num_anchors(level_hw)
    return tuple(like.new_empty((A, 4), dtype=_F32) for _ in range(3))


# --------------------------------------------------------------------------- a11
@torch.library.custom_op("sihl_b200::topk_locations", mutates_args=())
def topk_locations(loc_logits: Tensor, k: int) -> Tuple[Tensor, Tensor]:
    """ref :109 -> (values fp32 [B,k] (half logits are exact in fp32), indices [B,k] int64)."""
    return ops.topk_locations(loc_logits, k)


@topk_locations.register_fake
def _(loc_logits, k):
    B = loc_logits.shape[0]
    return loc_logits.new_empty((B, k), dtype=_F32), loc_logits.new_empty((B, k), dtype=_I64)


@torch.library.custom_op("sihl_b200::decode_rows", mutates_args=())
def decode_rows(top_logits: Tensor, idx: Tensor, cls_rows: Tensor, box_rows: Tensor, level_hw: List[int], img_w: int,
                img_h: int) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """ref :113-121 -> (num_instances i64 [B], scores f32 [B,K], classes i64 [B,K], boxes f32 [B,K,4] px)."""
    offsets, scales, _ = ops.anchor_tables(_pairs(level_hw), img_w, img_h, top_logits.device)
    return ops.decode_rows(top_logits, idx, cls_rows, box_rows, offsets, scales, img_w, img_h)


@decode_rows.register_fake
def _(top_logits, idx, cls_rows, box_rows, level_hw, img_w, img_h):
    B, K = top_logits.shape
    e = top_logits.new_empty
    return e((B,), dtype=_I64), e((B, K), dtype=_F32), e((B, K), dtype=_I64), e((B, K, 4), dtype=_F32)


# --------------------------------------------------------------------------- a15 (extension)
@torch.library.custom_op("sihl_b200::dense_postprocess", mutates_args=())
def dense_postprocess(loc_logits: Tensor, cls_logits: Tensor, box_raw: Tensor, level_hw: List[int], img_w: int, img_h: int,
                      score_thr: float, iou_thr: float, max_instances: int) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """Dense decode + class-aware NMS, output format of ``forward``."""
    return ops.dense_postprocess(loc_logits, cls_logits, box_raw, _pairs(level_hw), img_w, img_h, score_thr, iou_thr,
                                 max_instances)


@dense_postprocess.register_fake
def _(loc_logits, cls_logits, box_raw, level_hw, img_w, img_h, score_thr, iou_thr, max_instances):
    B, K = loc_logits.shape[0], max_instances
    e = loc_logits.new_empty
    return e((B,), dtype=_I64), e((B, K), dtype=_F32), e((B, K), dtype=_I64), e((B, K, 4), dtype=_F32)


# --------------------------------------------------------------------------- training step (N2)
@torch.library.custom_op("sihl_b200::train_assign", mutates_args=())
def train_assign(gt_boxes: Tensor, gt_offsets: Tensor, level_hw: List[int], img_w: int, img_h: int, topk: int,
                 pos_capacity: int) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """ref :134-148 + :182-184 -> (assignment i64 [B,A], rel_iou f32 [B,A], pos_index i32 [capacity], meta i32
    [16+B+2] = sums f64[8] | gt_offsets | pos_total).  ``gt_offsets`` int32 [B+1] on the device."""
    B = gt_offsets.shape[0] - 1
    st = ops.train_assign(_pairs(level_hw), img_w, img_h, gt_boxes, None, None, B, topk, gt_offsets=gt_offsets,
                          pos_capacity=pos_capacity)
    return st.assignment, st.rel_iou, st.pos_index, st.meta


@train_assign.register_fake
def _(gt_boxes, gt_offsets, level_hw, img_w, img_h, topk, pos_capacity):
    B, A = gt_offsets.shape[0] - 1, _num_anchors(level_hw)
    e = gt_boxes.new_empty
    return e((B, A), dtype=_I64), e((B, A), dtype=_F32), e((pos_capacity,), dtype=_I32), e((16 + B + 2,), dtype=_I32)


def _state(assignment, rel_iou, pos_index, meta, gt_boxes, gt_classes, level_hw, img_w, img_h, grad_scale=1.0):
    offsets, scales, _ = ops.anchor_tables(_pairs(level_hw), img_w, img_h, rel_iou.device)
    return ops.TrainAssignment.from_tensors(assignment, rel_iou, pos_index, meta, offsets, scales, gt_boxes, gt_classes,
                                            img_w, img_h, grad_scale)


@torch.library.custom_op("sihl_b200::train_loss", mutates_args=())
def train_loss(loc_logits: Tensor, iou_preds: Tensor, box_rows: Tensor, cls_rows: Tensor, assignment: Tensor,
               rel_iou: Tensor, pos_index: Tensor, meta: Tensor, gt_boxes: Tensor, gt_classes: Tensor,
               level_hw: List[int], img_w: int, img_h: int) -> Tuple[Tensor, Tensor]:
    """ref :157-217 -> (losses f32 [5] = [location, box, class, iou, total], meta' = a copy of ``meta`` whose 8 sums are
    filled in: the op is functional, the backward reads the normalisers from meta')."""
    meta = meta.clone()
    st = _state(assignment, rel_iou, pos_index, meta, gt_boxes, gt_classes, level_hw, img_w, img_h)
    losses, _ = ops.train_loss(st, loc_logits, iou_preds, box_rows, cls_rows, finalize=True)
    return losses, meta


@train_loss.register_fake
def _(loc_logits, iou_preds, box_rows, cls_rows, assignment, rel_iou, pos_index, meta, gt_boxes, gt_classes, level_hw,
      img_w, img_h):
    return loc_logits.new_empty((5,), dtype=_F32), torch.empty_like(meta)


@torch.library.custom_op("sihl_b200::train_loss_bwd", mutates_args=())
def train_loss_bwd(grad_losses: Tensor, loc_logits: Tensor, iou_preds: Tensor, box_rows: Tensor, cls_rows: Tensor,
                   assignment: Tensor, rel_iou: Tensor, pos_index: Tensor, meta: Tensor, gt_boxes: Tensor,
                   gt_classes: Tensor, level_hw: List[int], img_w: int, img_h: int) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    st = _state(assignment, rel_iou, pos_index, meta, gt_boxes, gt_classes, level_hw, img_w, img_h)
    maps, _, _ = ops._train_maps(st, loc_logits, iou_preds, box_rows, cls_rows)
    g = ops.train_loss_bwd(st, maps, grad_losses.float().contiguous())
    return tuple(gi.to(m.dtype) for gi, m in zip(g, (loc_logits, iou_preds, box_rows, cls_rows)))


@train_loss_bwd.register_fake
def _(grad_losses, loc_logits, iou_preds, box_rows, cls_rows, *rest):
    return tuple(torch.empty_like(t) for t in (loc_logits, iou_preds, box_rows, cls_rows))


def _train_loss_setup(ctx, inputs, output):
    t = list(inputs[:10])
    t[7] = output[1]                        # the meta block with this step's sums
    ctx.save_for_backward(*t)
    ctx.rest = inputs[10:]


def _train_loss_backward(ctx, grad, grad_meta):
    t = ctx.saved_tensors
    g = train_loss_bwd(grad, *t, *ctx.rest)
    return (g[0], g[1], g[2], g[3]) + (None,) * 9


train_loss.register_autograd(_train_loss_backward, setup_context=_train_loss_setup)


# --------------------------------------------------------------------------- a3 (static API, one image)
@torch.library.custom_op("sihl_b200::bbox_matching", mutates_args=())
def bbox_matching(anchors: Tensor, gt_boxes: Tensor, topk: int, relative: bool) -> Tuple[Tensor, Tensor]:
    """ref :252-284 for arbitrary anchors -> (assignment i64 [A], iou f32 [A])."""
    a, v = ops.bbox_matching(anchors, gt_boxes, topk, relative)
    return a.clone(), v.clone()


@bbox_matching.register_fake
def _(anchors, gt_boxes, topk, relative):
    A = anchors.shape[0]
    return anchors.new_empty((A,), dtype=_I64), anchors.new_empty((A,), dtype=_F32)


REGISTERED = ("anchor_tables", "topk_locations", "decode_rows", "dense_postprocess", "train_assign", "train_loss",
              "train_loss_bwd", "bbox_matching")

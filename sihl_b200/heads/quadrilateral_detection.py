"""The assigner of ``sihl.heads.QuadrilateralDetection`` on the B200 kernels.

Only the dense, non-learned assignment of that head is in scope: ``bbox_matching`` (ref
``src/sihl/heads/quadrilateral_detection.py:266-294``), the per-image loop around it (:165-172) and the anchor
construction that feeds it (:92-108, :156-163).  The learned part of the head and its L1 / focal losses stay with
the reference.  ``InstanceSegmentation`` (ref instance_segmentation.py:196) and ``KeypointDetection``
(keypoint_detection.py:208) call ``ObjectDetection.bbox_matching(..., relative=True)`` verbatim and are served by
:func:`sihl_b200.heads.ObjectDetection.bbox_matching`.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
from torch import Tensor

from .. import ops


def quads_to_boxes(quads: Tensor) -> Tensor:
    """ref :318-324 — axis-aligned bounds ``[N,4]`` (xyxy) of quads ``[N,4,2]``."""
    x, y = quads[..., 0], quads[..., 1]
    return torch.stack([x.min(-1).values, y.min(-1).values, x.max(-1).values, y.max(-1).values], 1)


def quad_anchors(level_sizes: Sequence[Tuple[int, int]], levels: Sequence[int], top_level: int, img_w: int, img_h: int,
                 device) -> Tensor:
    """ref :92-108 + :156-163 — one square anchor per location, half-extent ``sigmoid(level - top_level)`` of the
    image, centred on the cell centre: ``(rel_offsets + [-1,-1,1,1] * scale) * [W,H,W,H]``.  Same torch operators
    in the same order as the reference (a function of shapes only; callers cache it)."""
    rel, lvl = [], []
    for (h, w), level in zip(level_sizes, levels):
        y_min, x_min = 1 / h / 2, 1 / w / 2
        ys = torch.linspace(y_min, 1 - y_min, steps=h, device=device)
        xs = torch.linspace(x_min, 1 - x_min, steps=w, device=device)
        grid = torch.stack([xs.view(1, w).expand(h, w), ys.view(h, 1).expand(h, w)], dim=2).reshape(h * w, 2)
        rel.append(grid)
        lvl.append(torch.full((h * w, 1), level, device=device))
    rel_offsets, lvls = torch.cat(rel).repeat(1, 2), torch.cat(lvl)
    directions = torch.tensor([[-1, -1, 1, 1]], device=device)
    scale = torch.sigmoid(lvls - top_level)
    return ((rel_offsets + directions * scale) * torch.tensor([[img_w, img_h] * 2], device=device)).contiguous()


def batched_bbox_matching(anchors: Tensor, boxes: List[Tensor], topk: int) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """ref :165-172 — the per-image loop and the four stacks in two kernel launches:
    ``(assignment i64 [B,A], o2o_mask bool [B,A], o2m_iou f32 [B,A], rel_iou f32 [B,A])``."""
    gt = ops.GtBatch.from_lists(boxes, None, anchors.device)
    out = ops.quad_bbox_matching(anchors.float(), gt, topk)
    return out["assignment"], out["o2o_mask"], out["iou"], out["rel_iou"]


def bbox_matching(anchors: Tensor, gt_boxes: Tensor, topk: int) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """Drop-in for the static ``QuadrilateralDetection.bbox_matching`` (ref :266-294), one image.  ``assignment`` is
    canonical: -1 wherever ``rel_iou`` is not > 0 (the reference keeps the index of an arbitrary zero entry there and
    only ever reads ``assignment[rel_iou > 0]``, ref :188,:201)."""
    a, o, i, r = batched_bbox_matching(anchors, [gt_boxes.as_subclass(Tensor).reshape(-1, 4)], topk)
    return a[0], o[0], i[0], r[0]

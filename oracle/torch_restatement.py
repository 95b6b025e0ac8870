"""torch/torchvision restatement of the reference head's dense tail.

TEST INFRASTRUCTURE ONLY (same rules as ``od_oracle.c``).

Why a second oracle: "bit-exact against the reference" is only well defined
against the reference *run on the same device* — ``atan``/``linspace``/``exp``
differ in the last bit between torch-CPU (Sleef), glibc and CUDA libdevice
(SURVEY.md §7.1).  This file states the same computation as the reference with
the same torch / torchvision operators (torchvision is a dependency of the
reference that is installed on the GPU box; the reference source is not), so
the ``gpu`` tests can run it on ``cuda:0`` next to the kernels and demand bit
equality, while ``tests/test_restatement_vs_reference.py`` (container only)
proves it equal to the real reference functions on CPU.

Reference lines: ``src/sihl/heads/object_detection.py`` (cited per function).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
from torch import Tensor
from torch.nn import functional as F
from torchvision import ops as tvops
from torchvision.ops import boxes as tvboxes


def offsets_and_scales(levels: Sequence[Tuple[int, int]], device) -> Tuple[Tensor, Tensor]:
    """ref :83-97 — normalised cell centres and half-cell extents, level-major, row-major."""
    all_offsets, all_scales = [], []
    for h, w in levels:
        half_y, half_x = 1 / h / 2, 1 / w / 2
        ys = torch.linspace(half_y, 1 - half_y, steps=h, device=device)
        xs = torch.linspace(half_x, 1 - half_x, steps=w, device=device)
        gx = xs.view(1, w).expand(h, w)
        gy = ys.view(h, 1).expand(h, w)
        all_offsets.append(torch.stack([gx, gy, gx, gy], dim=2).reshape(h * w, 4))
        cell = torch.tensor([-half_x, -half_y, half_x, half_y], device=device)
        all_scales.append(cell.view(1, 4).expand(h * w, 4))
    return torch.cat(all_offsets), torch.cat(all_scales)


def full_size(img_w: int, img_h: int, device) -> Tensor:
    """ref :134-136 — an int64 row that promotes to fp32 in the products."""
    return torch.tensor([[img_w, img_h, img_w, img_h]], device=device)


def anchors_px(levels, img_w: int, img_h: int, device) -> Tensor:
    """ref :139-140."""
    off, sc = offsets_and_scales(levels, device)
    return (off + sc) * full_size(img_w, img_h, device)


def match_one(anchors: Tensor, gt: Tensor, topk: int = 9, relative: bool = True) -> Tuple[Tensor, Tensor]:
    """ref :252-284, raw output (zero-IoU fill slots as torch.topk leaves them)."""
    A, G = anchors.shape[0], gt.shape[0]
    dev = anchors.device
    assign = torch.full((A,), -1, device=dev)
    out = torch.zeros((A,), device=dev)
    if G == 0:
        return assign, out
    ciou = tvops.complete_box_iou(anchors, gt).clamp(0)                     # :263
    top_val, top_idx = torch.topk(ciou, k=topk, dim=0)                      # :264
    chosen = torch.zeros((A, G), dtype=torch.bool, device=dev)
    chosen.scatter_(0, top_idx, True)                                       # :267-268
    row_max, row_arg = torch.max(ciou * chosen.float(), dim=1)              # :270
    hit = chosen.any(dim=1)                                                 # :271
    assign[hit] = row_arg[hit]                                              # :272
    if relative:
        denom = top_val[0][row_arg]                                         # :277-278
        out[hit] = (row_max[hit] / denom[hit]).nan_to_num(0)                # :279-281
    else:
        out[hit] = row_max[hit]                                             # :273
    return assign, out


def quad_match_one(anchors: Tensor, gt: Tensor, topk: int = 9) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """ref quadrilateral_detection.py:266-294, raw output (operator for operator)."""
    A, G = anchors.shape[0], gt.shape[0]
    dev = anchors.device
    assign = torch.full((A,), -1, device=dev)                               # :271
    o2o = torch.zeros((A,), dtype=torch.bool, device=dev)
    iou = torch.zeros((A,), device=dev)
    rel = torch.zeros((A,), device=dev)
    if G == 0:
        return assign, o2o, iou, rel
    ciou = tvops.complete_box_iou(anchors, gt)                              # :277 (no clamp)
    top_val, top_idx = torch.topk(ciou, k=topk, dim=0)                      # :278
    best_match = torch.zeros((A, G), dtype=torch.bool, device=dev)
    best_match.scatter_(0, top_idx[0:1], True)                              # :279-280
    chosen = torch.zeros((A, G), dtype=torch.bool, device=dev)
    chosen.scatter_(0, top_idx, True)                                       # :281-282
    row_max, row_arg = torch.max(ciou * chosen.float(), dim=1)              # :283
    hit = chosen.any(dim=1)                                                 # :284
    assign[hit] = row_arg[hit]                                              # :285
    o2o = best_match.any(dim=1)                                             # :286
    iou[hit] = row_max[hit]                                                 # :287
    denom = top_val[0][row_arg]                                             # :289-290
    rel[hit] = (row_max[hit] / denom[hit]).nan_to_num(0)                    # :291-293
    return assign, o2o, iou, rel


def quad_canonical(assign: Tensor, o2o: Tensor, iou: Tensor, rel: Tensor):
    """Canonical form of the four outputs: assignment defined only where rel_iou > 0; signed zeros folded."""
    assign = assign.clone()
    assign[~(rel > 0)] = -1
    return assign, o2o, iou + 0.0, rel + 0.0


def canonical(assign: Tensor, iou: Tensor) -> Tuple[Tensor, Tensor]:
    """SURVEY.md §3.4 contract: assignment is defined only where iou > 0."""
    assign = assign.clone()
    assign[~(iou > 0)] = -1
    return assign, iou


def assign_batch(anchors: Tensor, boxes: List[Tensor], topk: int = 9) -> Tuple[Tensor, Tensor]:
    """ref :143-148."""
    res = [match_one(anchors, b, topk, relative=True) for b in boxes]
    return torch.stack([r[0] for r in res]), torch.stack([r[1] for r in res])


def train_losses(levels, img_w: int, img_h: int, boxes: List[Tensor], classes: List[Tensor],
                 loc_logits: Tensor, iou_preds: Tensor, box_raw: Tensor, cls_logits: Tensor,
                 topk: int = 9):
    """ref :134-217 with the four MLP heads replaced by dense synthetic maps
    (``loc_logits``/``iou_preds`` [B,A]; ``box_raw`` [B,A,4], ``cls_logits`` [B,A,C]
    indexed at the positives — what the heads would have produced for those rows).
    Returns (loss, metrics dict, assignment, rel_iou)."""
    dev = loc_logits.device
    size = full_size(img_w, img_h, dev)
    offsets, scales = offsets_and_scales(levels, dev)
    anchors = (offsets + scales) * size
    assignment, rel = assign_batch(anchors, boxes, topk)

    one_hot = (rel == 1.0).to(torch.float32)                                               # :159
    loc_loss = F.binary_cross_entropy_with_logits(loc_logits, one_hot, reduction="none")
    loc_loss = loc_loss.sum() / one_hot.sum()                                              # :163
    if rel.max() == 0:                                                                     # :165-172
        z = torch.zeros_like(loc_loss)
        return loc_loss, dict(location_loss=loc_loss, box_loss=z, class_loss=z, iou_loss=z), assignment, rel

    iou_loss = F.mse_loss(iou_preds.to(torch.float32), rel, reduction="none").sum() / rel.sum()   # :177-180
    pos = rel > 0                                                                          # :182
    weight = rel[pos]
    pos_off = torch.cat([offsets[m] for m in pos])                                         # :187-188
    pos_sc = torch.cat([scales[m] for m in pos])
    pred = pos_off + pos_sc * box_raw[pos].exp()                                           # :189
    tgt_box = torch.cat([boxes[b][assignment[b, m]] for b, m in enumerate(pos)])           # :190-192
    box_loss = tvops.complete_box_iou_loss(pred, tgt_box.to(torch.float32) / size, reduction="none")
    box_loss = (weight * box_loss).sum() / weight.sum()                                    # :197
    tgt_cls = torch.cat([classes[b][assignment[b, m]] for b, m in enumerate(pos)])         # :201-203
    cls_loss = F.cross_entropy(cls_logits[pos].to(torch.float32), tgt_cls, reduction="none")
    cls_loss = (weight * cls_loss).sum() / weight.sum()                                    # :208
    loss = loc_loss + 10 * box_loss + cls_loss + iou_loss                                  # :210
    return loss, dict(location_loss=loc_loss, box_loss=box_loss, class_loss=cls_loss, iou_loss=iou_loss), assignment, rel


def forward_tail(levels, img_w: int, img_h: int, loc_logits: Tensor, box_raw: Tensor, cls_logits: Tensor,
                 max_instances: int = 100):
    """ref :106-122 on dense maps: top-k locations, sigmoid scores, argmax class, decoded boxes."""
    dev = loc_logits.device
    B = loc_logits.shape[0]
    offsets, scales = offsets_and_scales(levels, dev)
    top, idx = loc_logits.topk(max_instances, dim=1)                                       # :109
    rows = torch.arange(B, device=dev).view(B, 1).expand(B, max_instances)
    scores = top.sigmoid()                                                                 # :113
    num = (scores > 0.5).sum(dim=1)                                                        # :114
    cls = cls_logits[rows, idx].max(dim=2).indices                                         # :116-117
    off, sc = offsets[idx], scales[idx]                                                    # :119-120
    boxes = (off + sc * box_raw[rows, idx].exp()) * full_size(img_w, img_h, dev).view(1, 1, 4)   # :121
    return num, scores, cls, boxes, idx


def nms_per_class(boxes: Tensor, scores: Tensor, classes: Tensor, iou_thr: float) -> Tensor:
    """Extension oracle: torchvision's exact per-class path (tv:ops/boxes.py:102-120)."""
    if boxes.numel() == 0:
        return torch.empty((0,), dtype=torch.int64, device=boxes.device)
    return tvboxes._batched_nms_vanilla(boxes, scores, classes, iou_thr)


def dense_postprocess(levels, img_w: int, img_h: int, loc_logits: Tensor, box_raw: Tensor, cls_logits: Tensor,
                      score_thr: float = 0.05, iou_thr: float = 0.5, max_instances: int = 100):
    """Extension: decode every location (score = sigmoid(loc), class = argmax), keep
    score > thr, class-aware NMS, first ``max_instances`` by score; zero padded."""
    dev = loc_logits.device
    B, A = loc_logits.shape
    K = max_instances
    offsets, scales = offsets_and_scales(levels, dev)
    size = full_size(img_w, img_h, dev)
    num = torch.zeros(B, dtype=torch.int64, device=dev)
    scores = torch.zeros((B, K), device=dev)
    classes = torch.zeros((B, K), dtype=torch.int64, device=dev)
    boxes = torch.zeros((B, K, 4), device=dev)
    for b in range(B):
        s = loc_logits[b].sigmoid()
        cand = (s > score_thr).nonzero().squeeze(1)
        cb = (offsets[cand] + scales[cand] * box_raw[b, cand].exp()) * size
        cc = cls_logits[b, cand].max(dim=1).indices if cand.numel() else cand
        keep = nms_per_class(cb, s[cand], cc, iou_thr)[:K]
        m = keep.numel()
        num[b] = m
        scores[b, :m], classes[b, :m], boxes[b, :m] = s[cand][keep], cc[keep], cb[keep]
    return num, scores, classes, boxes

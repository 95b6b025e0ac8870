"""Synthetic inputs for the ObjectDetection-head hot path.

Two generators with the same distributions:

* ``numpy`` (``np.random.RandomState`` — a frozen bit stream, so the committed
  golden fixtures can store a seed instead of the inputs) for tests and golden
  vectors;
* ``torch`` on a device (``torch.Generator``) for ``bench.py``.

Distributions: level sizes from the image size, gt box sides
log-uniform in ``[8, 0.625*S]`` px with uniform centres, clamped to the image,
redrawn when a side collapses below 1 px, non-integer coordinates;
``loc_logits ~ N(-5, 1)`` (the head's initial bias, reference
``object_detection.py:58``), ``iou_preds ~ U(0,1)``, ``box_raw ~ N(0, 0.5^2)``,
``cls_logits ~ N(0, 1)``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np


def level_sizes(height: int, width: int, bottom_level: int = 3, top_level: int = 7,
                mode: str = "ceil") -> List[Tuple[int, int]]:
    """Feature-map sizes ``(h_l, w_l)`` for levels ``bottom..top``.

    ``mode="ceil"`` halves with ceil at every level (what stride-2 convs with
    padding produce: 320 -> 160, 80, 40, 20, 10, 5, 3), ``mode="floor"`` uses
    ``S // 2**l``.  The head itself never assumes either: it reads the sizes
    off the tensors it is given (reference ``object_detection.py:87``).
    """
    sizes = []
    for level in range(bottom_level, top_level + 1):
        if mode == "floor":
            sizes.append((max(height // 2 ** level, 1), max(width // 2 ** level, 1)))
        else:
            h, w = height, width
            for _ in range(level):
                h, w = (h + 1) // 2, (w + 1) // 2
            sizes.append((h, w))
    return sizes


def num_anchors(levels: Sequence[Tuple[int, int]]) -> int:
    return int(sum(h * w for h, w in levels))


def gt_boxes_np(rng: np.random.RandomState, n: int, height: int, width: int,
                min_side: float = 8.0, max_frac: float = 0.625) -> np.ndarray:
    """``n`` xyxy pixel boxes, fp32, non-lattice coordinates."""
    out = np.zeros((n, 4), dtype=np.float32)
    s = float(min(height, width))
    lo, hi = math.log(min_side), math.log(max(max_frac * s, min_side * 1.5))
    i = 0
    while i < n:
        bw, bh = np.exp(rng.uniform(lo, hi, size=2))
        cx, cy = rng.uniform(0, width), rng.uniform(0, height)
        x1, y1 = max(cx - bw / 2, 0.0), max(cy - bh / 2, 0.0)
        x2, y2 = min(cx + bw / 2, float(width)), min(cy + bh / 2, float(height))
        box = np.array([x1, y1, x2, y2], dtype=np.float32)
        if box[2] - box[0] < 1.0 or box[3] - box[1] < 1.0:
            continue
        out[i] = box
        i += 1
    return out


@dataclass
class GtBatch:
    """Ragged ground truth in CSR form (what the C-ABI consumes)."""
    boxes: np.ndarray      # [sumG, 4] f32 xyxy px
    classes: np.ndarray    # [sumG] i64
    offsets: np.ndarray    # [B+1] i32

    @property
    def batch_size(self) -> int:
        return len(self.offsets) - 1

    def per_image(self):
        for b in range(self.batch_size):
            s, e = int(self.offsets[b]), int(self.offsets[b + 1])
            yield self.boxes[s:e], self.classes[s:e]


def gt_batch_np(seed: int, batch: int, height: int, width: int, num_classes: int,
                max_gt: int, ragged: bool = True, counts: Optional[Sequence[int]] = None,
                integer_coords: bool = False) -> GtBatch:
    """Ragged (``G_b ~ U{0..max_gt}``, image 0 empty when ``batch>1``) or full."""
    rng = np.random.RandomState(seed)
    if counts is None:
        if ragged:
            counts = [int(rng.randint(0, max_gt + 1)) for _ in range(batch)]
            if batch > 1:
                counts[0] = 0
            if batch > 2:
                counts[1] = max_gt
        else:
            counts = [max_gt] * batch
    boxes, classes = [], []
    for n in counts:
        b = gt_boxes_np(rng, n, height, width)
        if integer_coords:
            b = np.round(b).astype(np.float32)
        boxes.append(b)
        classes.append(rng.randint(0, num_classes, size=n).astype(np.int64))
    offsets = np.zeros(len(counts) + 1, dtype=np.int32)
    offsets[1:] = np.cumsum(counts)
    return GtBatch(np.concatenate(boxes).reshape(-1, 4).astype(np.float32),
                   np.concatenate(classes).astype(np.int64), offsets)


@dataclass
class DenseMaps:
    loc_logits: np.ndarray   # [B, A]
    iou_preds: np.ndarray    # [B, A]
    box_raw: np.ndarray      # [B, A, 4]
    cls_logits: np.ndarray   # [B, A, C]


def dense_maps_np(seed: int, batch: int, anchors: int, num_classes: int,
                  loc_mean: float = -5.0, loc_std: float = 1.0) -> DenseMaps:
    rng = np.random.RandomState(seed)
    f32 = np.float32
    return DenseMaps(
        loc_logits=(loc_mean + loc_std * rng.standard_normal((batch, anchors))).astype(f32),
        iou_preds=rng.uniform(0, 1, (batch, anchors)).astype(f32),
        box_raw=(0.5 * rng.standard_normal((batch, anchors, 4))).astype(f32),
        cls_logits=rng.standard_normal((batch, anchors, num_classes)).astype(f32),
    )


def nms_candidates_np(seed: int, n: int, size: int, num_classes: int = 80):
    """NMS stress inputs (SURVEY.md §8d config[3]): sides U(4, 0.3*S), scores U(0.05,1)."""
    rng = np.random.RandomState(seed)
    w = rng.uniform(4, 0.3 * size, n)
    h = rng.uniform(4, 0.3 * size, n)
    cx, cy = rng.uniform(0, size, n), rng.uniform(0, size, n)
    boxes = np.stack([np.clip(cx - w / 2, 0, size), np.clip(cy - h / 2, 0, size),
                      np.clip(cx + w / 2, 0, size), np.clip(cy + h / 2, 0, size)], 1).astype(np.float32)
    bad = (boxes[:, 2] - boxes[:, 0] < 1) | (boxes[:, 3] - boxes[:, 1] < 1)
    boxes[bad] = np.array([0, 0, 8, 8], np.float32) + rng.uniform(0, size - 8, (int(bad.sum()), 1)).astype(np.float32)
    scores = rng.uniform(0.05, 1.0, n).astype(np.float32)
    classes = rng.randint(0, num_classes, n).astype(np.int64)
    return boxes, scores, classes


# ----------------------------------------------------------------------------
# device-side generator for bench.py (torch imported lazily: synth is also used
# by numpy-only code)
# ----------------------------------------------------------------------------
def gt_batch_torch(gen, batch: int, height: int, width: int, num_classes: int, max_gt: int,
                   device, min_side: float = 8.0, max_frac: float = 0.625):
    """Full (``G_b == max_gt``) gt batch generated on ``device``.

    Returns ``(boxes [B*G,4] f32, classes [B*G] i64, offsets [B+1] i32)``.
    Boxes whose side collapses under 1 px after clamping are replaced by a
    centred box of the minimum side (the numpy generator redraws instead).
    """
    import torch
    n = batch * max_gt
    s = float(min(height, width))
    lo, hi = math.log(min_side), math.log(max(max_frac * s, min_side * 1.5))
    u = torch.rand((n, 4), generator=gen, device=device, dtype=torch.float64)
    bw = torch.exp(lo + (hi - lo) * u[:, 0])
    bh = torch.exp(lo + (hi - lo) * u[:, 1])
    cx, cy = u[:, 2] * width, u[:, 3] * height
    x1, y1 = (cx - bw / 2).clamp(min=0), (cy - bh / 2).clamp(min=0)
    x2, y2 = (cx + bw / 2).clamp(max=width), (cy + bh / 2).clamp(max=height)
    bad = ((x2 - x1) < 1) | ((y2 - y1) < 1)
    x1 = torch.where(bad, torch.full_like(x1, width / 2 - min_side / 2 + 0.37), x1)
    x2 = torch.where(bad, torch.full_like(x2, width / 2 + min_side / 2 + 0.37), x2)
    y1 = torch.where(bad, torch.full_like(y1, height / 2 - min_side / 2 + 0.41), y1)
    y2 = torch.where(bad, torch.full_like(y2, height / 2 + min_side / 2 + 0.41), y2)
    boxes = torch.stack([x1, y1, x2, y2], 1).to(torch.float32).contiguous()
    classes = torch.randint(0, num_classes, (n,), generator=gen, device=device, dtype=torch.int64)
    offsets = (torch.arange(batch + 1, device=device, dtype=torch.int32) * max_gt).contiguous()
    return boxes, classes, offsets


def dense_maps_torch(gen, batch: int, anchors: int, num_classes: int, device,
                     loc_mean: float = -5.0, loc_std: float = 1.0):
    import torch
    f32 = torch.float32
    loc = torch.randn((batch, anchors), generator=gen, device=device, dtype=f32) * loc_std + loc_mean
    iou = torch.rand((batch, anchors), generator=gen, device=device, dtype=f32)
    box = torch.randn((batch, anchors, 4), generator=gen, device=device, dtype=f32) * 0.5
    cls = torch.randn((batch, anchors, num_classes), generator=gen, device=device, dtype=f32)
    return loc, iou, box, cls

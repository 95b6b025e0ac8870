"""The hot path executed by the REAL reference code on synthetic head outputs — the baseline arm of ``bench.py``.

TEST INFRASTRUCTURE ONLY (bench.py's ``--impl reference`` / ``cpu_baseline`` / ``gpu_eager_reference`` legs,
``oracle/make_golden.py``); nothing under ``sihl_b200/`` imports it.

``ReferencePath`` wraps the reference's unmodified ``ObjectDetection`` head (``oracle/ref_loader.py``:
/root/reference or the staged copy) whose four MLPs are replaced by table look-ups into the synthetic dense maps and
whose laterals are identities — so ``training_step`` (ref object_detection.py:124-217: anchors, the per-image
``bbox_matching`` loop, compaction, the four losses with their torchvision CIoU calls) runs exactly as shipped, on
inputs we control, on any device.  The NMS leg is the north-star extension (the reference has none): every location
decoded with the semantics of ref :113-121, thresholded, then ``torchvision.ops.batched_nms`` — torchvision's own
dispatch (coordinate trick below 4000 boxes on CPU / 100000 on CUDA, per-class loop above).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
from torch import Tensor, nn
from torchvision import ops as tvops

from . import ref_loader


class _Table(nn.Module):
    """Stands in for an MLP head: the row id rides in channel 0 of the features."""

    def __init__(self, table: Tensor):
        super().__init__()
        self.table = nn.Parameter(table)

    def forward(self, feats):
        return self.table[feats[..., 0].round().long()]


def reference_head(levels: Sequence[Tuple[int, int]], height: int, width: int, bottom: int, top: int, num_classes: int,
                   batch: int, loc: Tensor, iou: Tensor, box: Tensor, cls: Tensor, max_instances: int = 100):
    """(head, inputs): the reference head over table look-ups + the level inputs that carry the row ids.
    ``loc``/``iou`` [B,A], ``box`` [B,A,4], ``cls`` [B,A,C] on the device the head should run on."""
    RefOD = ref_loader.ObjectDetection()
    dev = loc.device
    A = sum(h * w for h, w in levels)
    assert batch * A < (1 << 24), "row ids must stay exact in fp32"
    head = RefOD(in_channels=[3] + [4] * top, num_classes=num_classes, bottom_level=bottom, top_level=top,
                 num_channels=4, num_layers=0, max_instances=max_instances)
    head.laterals = nn.ModuleList([nn.Identity() for _ in levels])
    head.loc_head = _Table(loc.reshape(batch * A, 1).clone())
    head.iou_head = _Table(iou.reshape(batch * A, 1).clone())
    head.box_head = _Table(box.reshape(batch * A, 4).clone())
    head.cls_head = _Table(cls.reshape(batch * A, num_classes).clone())
    inputs = [torch.zeros(batch, 3, height, width, device=dev)] + [torch.zeros(batch, 1, 1, 1, device=dev) for _ in range(1, bottom)]
    start = 0
    for (h, w) in levels:
        ids = torch.arange(start, start + h * w, dtype=torch.float32, device=dev).view(1, 1, h, w)
        ids = ids + (torch.arange(batch, dtype=torch.float32, device=dev) * A).view(batch, 1, 1, 1)
        inputs.append(torch.cat([ids, torch.zeros(batch, 3, h, w, device=dev)], dim=1))
        start += h * w
    return head.to(dev), inputs


class ReferencePath:
    """One pass = ``training_step`` of the real reference head + dense decode + ``torchvision.ops.batched_nms``."""

    def __init__(self, levels, height: int, width: int, num_classes: int, boxes: List[Tensor], classes: List[Tensor],
                 loc: Tensor, iou: Tensor, box: Tensor, cls: Tensor, max_instances: int = 100, bottom: int = 3):
        self.levels, self.H, self.W, self.K = list(levels), int(height), int(width), int(max_instances)
        self.B = int(loc.shape[0])
        top = bottom + len(self.levels) - 1
        self.head, self.inputs = reference_head(self.levels, height, width, bottom, top, num_classes, self.B, loc, iou, box,
                                                cls, max_instances)
        self.boxes, self.classes = boxes, classes
        self.loc, self.box, self.cls = loc, box, cls

    @torch.no_grad()
    def train(self):
        """ref :124-217, unmodified (loss, metrics)."""
        return self.head.training_step(self.inputs, self.classes, self.boxes)

    def train_with_backward(self):
        for p in self.head.parameters():
            p.grad = None
        loss, metrics = self.head.training_step(self.inputs, self.classes, self.boxes)
        loss.backward()
        return loss, metrics

    @torch.no_grad()
    def forward(self):
        """ref :99-122, unmodified."""
        return self.head.forward(self.inputs)

    @torch.no_grad()
    def postprocess(self, score_thr: float, iou_thr: float):
        """Extension leg: dense decode (ref :113-121 semantics over all locations) + torchvision batched_nms."""
        offsets, scales = self.head.get_offsets_and_scales(self.inputs)          # ref :83-97
        size = torch.tensor([[self.W, self.H, self.W, self.H]], device=self.loc.device)
        out = []
        for b in range(self.B):
            s = self.loc[b].sigmoid()
            cand = (s > score_thr).nonzero().squeeze(1)
            cb = (offsets[cand] + scales[cand] * self.box[b, cand].exp()) * size
            cc = self.cls[b, cand].max(dim=1).indices if cand.numel() else cand
            keep = tvops.batched_nms(cb, s[cand], cc, iou_thr)[: self.K]
            out.append((s[cand][keep], cc[keep], cb[keep]))
        return out

    def one_pass(self, score_thr: float, iou_thr: float):
        loss, _ = self.train()
        dets = self.postprocess(score_thr, iou_thr)
        return loss, dets

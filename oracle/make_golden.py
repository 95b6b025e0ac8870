"""Generate ``tests/golden/*.npz`` by running the UNMODIFIED reference head.

Run in the authoring container only (needs ``/root/reference``):

    python -m oracle.make_golden            # from the repo root

Inputs are regenerated from a seed by ``sihl_b200.synth`` (numpy
``RandomState``, a frozen stream), so fixtures hold seeds + reference outputs
(+ the reference's own anchor tables).  The four MLP heads of the reference
``ObjectDetection`` are replaced by table look-ups into synthetic dense maps and
the laterals by identities, so that ``training_step`` / ``forward`` — the
reference's own code, including its compaction, per-image gathers, torchvision
CIoU calls and autograd — run unmodified on inputs we control.

TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import torch
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from sihl_b200 import synth  # noqa: E402
from oracle.golden_cases import (ASSIGN_CASES, FORWARD_CASES, GEOMETRIES, GOLDEN_DIR, NMS_CASES, QUAD_CASES,  # noqa: E402
                                 TRAIN_CASES, case_gt, forward_maps, geom_levels, nms_inputs, quad_gt, train_maps)

def reference_head(levels, g, num_classes, batch, maps: synth.DenseMaps, max_instances=100):
    """The unmodified reference head over table look-ups (oracle/ref_path.py)."""
    from oracle import ref_path
    t = torch.from_numpy
    return ref_path.reference_head(levels, g["height"], g["width"], g["bottom"], g["top"], num_classes, batch,
                                   t(maps.loc_logits), t(maps.iou_preds), t(maps.box_raw), t(maps.cls_logits), max_instances)


def save(name, **arrays):
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    path = os.path.join(GOLDEN_DIR, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name:24s} {os.path.getsize(path) / 1024:8.1f} KiB")


def make_geometry(name, g):
    RefOD = ref_loader.ObjectDetection()
    levels = geom_levels(g)
    head = RefOD(in_channels=[3] + [4] * g["top"], num_classes=1, bottom_level=g["bottom"], top_level=g["top"],
                 num_channels=4, num_layers=0)
    inputs = [torch.zeros(1, 1, g["height"], g["width"])] + [None] * (g["bottom"] - 1) + \
             [torch.zeros(1, 1, h, w) for h, w in levels]
    off, sc = head.get_offsets_and_scales(inputs)                                   # ref :83-97
    size = torch.tensor([[g["width"], g["height"], g["width"], g["height"]]])       # ref :134-136
    anchors = (off + sc) * size                                                     # ref :140
    save("geom_" + name, levels=np.asarray(levels, np.int32), img_wh=np.asarray([g["width"], g["height"]], np.int32),
         offsets=off.numpy(), scales=sc.numpy(), anchors=anchors.numpy())
    return anchors


def make_assign(name, case, anchors):
    RefOD = ref_loader.ObjectDetection()
    gt = case_gt(case)
    out = {}
    for relative in (True, False):
        flat, vals, asg, n_valid_raw = [], [], [], []
        for b, (bx, _) in enumerate(gt.per_image()):
            a, v = RefOD.bbox_matching(anchors, torch.from_numpy(bx), 9, relative=relative)   # ref :252-284
            a, v = a.numpy(), v.numpy()
            assert not np.isnan(v).any()
            n_valid_raw.append(int((a >= 0).sum()))
            pos = np.nonzero(v > 0)[0]
            flat.append(pos + b * len(a)); vals.append(v[pos]); asg.append(a[pos])
        tag = "rel" if relative else "abs"
        out[f"{tag}_pos_flat"] = np.concatenate(flat).astype(np.int32)
        out[f"{tag}_pos_val"] = np.concatenate(vals).astype(np.float32)
        out[f"{tag}_pos_assign"] = np.concatenate(asg).astype(np.int32)
        out[f"{tag}_n_valid_raw"] = np.asarray(n_valid_raw, np.int32)
    save(name, **out)


def make_train(name, case):
    g = GEOMETRIES[case["geom"]]
    levels = geom_levels(g)
    gt = case_gt(case)
    B, A, C = gt.batch_size, synth.num_anchors(levels), case["classes"]
    maps = train_maps(case)
    head, inputs = reference_head(levels, g, C, B, maps)
    boxes = [torch.from_numpy(bx) for bx, _ in gt.per_image()]
    classes = [torch.from_numpy(cl) for _, cl in gt.per_image()]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        loss, metrics = head.training_step(inputs, classes, boxes)                  # ref :124-217
    grads = {}
    if torch.isfinite(loss):
        loss.backward()
        for key, mod in (("dloc", head.loc_head), ("diou", head.iou_head), ("dbox", head.box_head), ("dcls", head.cls_head)):
            gr = mod.table.grad
            grads[key] = np.zeros(0, np.float32) if gr is None else gr.numpy()
        rows = np.nonzero(np.abs(grads["dcls"]).sum(1) + np.abs(grads["dbox"]).sum(1))[0] if grads["dcls"].size else np.zeros(0, np.int64)
        grads["grad_rows"] = rows.astype(np.int32)
        if grads["dcls"].size:
            grads["dcls"] = grads["dcls"][rows]
            grads["dbox"] = grads["dbox"][rows]
    save(name, loss=np.float64(loss.item()),
         location_loss=np.float64(metrics["location_loss"].item()), box_loss=np.float64(metrics["box_loss"].item()),
         class_loss=np.float64(metrics["class_loss"].item()), iou_loss=np.float64(metrics["iou_loss"].item()), **grads)


def make_forward(name, case):
    g = GEOMETRIES[case["geom"]]
    levels = geom_levels(g)
    B, A, C, K = case["batch"], synth.num_anchors(levels), case["classes"], case["k"]
    maps = forward_maps(case)
    head, inputs = reference_head(levels, g, C, B, maps, max_instances=K)
    with torch.no_grad():
        num, scores, classes, boxes = head.forward(inputs)                          # ref :99-122
        _, idx = torch.from_numpy(maps.loc_logits).topk(K, dim=1)                   # ref :109 (indices, for the parity test)
    save(name, num_instances=num.numpy(), scores=scores.numpy(), classes=classes.numpy().astype(np.int32),
         boxes=boxes.numpy(), idx=idx.numpy().astype(np.int32))


def make_nms(name, case):
    from torchvision.ops import boxes as tvb
    boxes, scores, classes = nms_inputs(case)
    tb, ts, tc = torch.from_numpy(boxes), torch.from_numpy(scores), torch.from_numpy(classes)
    keep = tvb._batched_nms_vanilla(tb, ts, tc, case["thr"]).numpy()                # tv:ops/boxes.py:102-120
    keep_trick = tvb._batched_nms_coordinate_trick(tb, ts, tc, case["thr"]).numpy() # tv:ops/boxes.py:87-100 (for information)
    save(name, keep=keep.astype(np.int32), keep_coordinate_trick=keep_trick.astype(np.int32))


def make_quad(name, case):
    """N1: the reference's QuadrilateralDetection — anchors as training_step builds them (ref :154-163, verbatim) and
    bbox_matching (ref :266-294) per image; outputs stored in canonical form (assignment only where rel_iou > 0)."""
    Q = ref_loader.QuadrilateralDetection()
    H, W, bottom, top = case["height"], case["width"], case["bottom"], case["top"]
    head = Q([3] + [8] * top, 4, bottom_level=bottom, top_level=top, num_channels=8, num_layers=1)
    sizes = [(H, W)] + [(-(-H // 2 ** l), -(-W // 2 ** l)) for l in range(1, top + 1)]
    inputs = [torch.zeros(1, 1, h, w) for h, w in sizes]
    rel_offsets, levels = head.get_offsets_and_levels(inputs)                                    # ref :156
    directions = torch.tensor([[-1, -1, 1, 1]])                                                  # ref :157
    scale = torch.sigmoid(levels - head.top_level)                                               # ref :158
    anchors = (rel_offsets[:, :4] + directions * scale) * torch.tensor([[W, H] * 2])             # ref :159-161
    gt = quad_gt(case)
    asg, o2o, iou, rel = [], [], [], []
    for bx, _ in gt.per_image():
        a, o, i, r = Q.bbox_matching(anchors, torch.from_numpy(bx).reshape(-1, 4), head.topk)    # ref :165-168
        assert not torch.isnan(i).any() and not torch.isnan(r).any()
        a = a.clone(); a[~(r > 0)] = -1
        asg.append(a.numpy()); o2o.append(o.numpy()); iou.append((i + 0.0).numpy()); rel.append((r + 0.0).numpy())
    save(name, anchors=anchors.numpy().astype(np.float32), level_sizes=np.asarray(sizes[bottom:top + 1], np.int32),
         assignment=np.stack(asg).astype(np.int64), o2o=np.stack(o2o), iou=np.stack(iou).astype(np.float32),
         rel=np.stack(rel).astype(np.float32), topk=np.int32(head.topk))


def main():
    if not ref_loader.available():
        raise SystemExit("reference tree not found; golden vectors can only be generated in the authoring container")
    torch.set_num_threads(4)
    anchors = {name: make_geometry(name, g) for name, g in GEOMETRIES.items()}
    for name, case in ASSIGN_CASES.items():
        make_assign(name, case, anchors[case["geom"]])
    for name, case in TRAIN_CASES.items():
        make_train(name, case)
    for name, case in FORWARD_CASES.items():
        make_forward(name, case)
    for name, case in NMS_CASES.items():
        make_nms(name, case)
    for name, case in QUAD_CASES.items():
        make_quad(name, case)


if __name__ == "__main__":
    main()

"""Case tables for the golden fixtures (shared by ``make_golden.py`` and ``tests/``).

TEST INFRASTRUCTURE ONLY.  Data + input regeneration; no reference access.
"""
from __future__ import annotations

import os

import numpy as np

from sihl_b200 import synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# name -> dict(height, width, bottom, top, mode, classes, counts|max_gt, seed, ...)
GEOMETRIES = {
    "cfg0_320": dict(height=320, width=320, bottom=3, top=7, mode="ceil"),       # resnet18+FPN at 320²: 40,20,10,5,3
    "test_128": dict(height=128, width=128, bottom=3, top=7, mode="floor"),      # reference test fixture: 16,8,4,2,1
    "cfg1_640": dict(height=640, width=640, bottom=3, top=7, mode="ceil"),       # 80,40,20,10,5 (A=8525)
    "nonsq_384x512": dict(height=384, width=512, bottom=3, top=7, mode="ceil"),
}

ASSIGN_CASES = {
    "assign_cfg0": dict(geom="cfg0_320", classes=10, counts=[20, 7], seed=101),
    "assign_test128": dict(geom="test_128", classes=16, counts=[0, 1, 2, 3], seed=102),
    "assign_cfg1": dict(geom="cfg1_640", classes=80, batch=4, max_gt=100, seed=103),
    "assign_nonsq": dict(geom="nonsq_384x512", classes=80, batch=3, max_gt=40, seed=104),
    "assign_ties": dict(geom="cfg1_640", classes=80, counts=[50, 50], seed=105, integer_coords=True),
    "assign_flipped": dict(geom="test_128", classes=16, counts=[6, 6], seed=106, flipped=True),
}

TRAIN_CASES = {
    "train_cfg0": dict(geom="cfg0_320", classes=10, counts=[20, 7], seed=201),
    "train_test128": dict(geom="test_128", classes=16, counts=[0, 1, 2, 3], seed=202),
    "train_cfg1": dict(geom="cfg1_640", classes=80, counts=[100, 37], seed=203),
    "train_empty": dict(geom="test_128", classes=16, counts=[0, 0], seed=204),
}

FORWARD_CASES = {
    "forward_cfg0": dict(geom="cfg0_320", classes=10, batch=2, seed=301, k=100),
    "forward_cfg1": dict(geom="cfg1_640", classes=80, batch=2, seed=302, k=100),
    "forward_test128": dict(geom="test_128", classes=16, batch=4, seed=303, k=100),
}

NMS_CASES = {
    "nms_small": dict(n=300, size=640, classes=80, seed=401, thr=0.5),
    "nms_mid": dict(n=3000, size=640, classes=80, seed=402, thr=0.5),
    "nms_fewclasses": dict(n=2000, size=1024, classes=3, seed=403, thr=0.45),
    "nms_oneclass": dict(n=700, size=640, classes=1, seed=404, thr=0.6),
}


# N1: QuadrilateralDetection.bbox_matching (levels 3..5 like the reference default; "tiny" = gts far smaller than every
# anchor, where each gt's best CIoU is negative and the un-clamped top-k / all-selected paths are exercised)
QUAD_CASES = {
    "quad_128": dict(height=128, width=128, bottom=3, top=5, counts=[0, 1, 3, 7], seed=501),
    "quad_320": dict(height=320, width=320, bottom=3, top=5, counts=[20, 5], seed=502),
    "quad_tiny": dict(height=128, width=160, bottom=3, top=5, counts=[1, 2, 9], seed=503, tiny=True),
}


def quad_gt(case) -> synth.GtBatch:
    gt = synth.gt_batch_np(case["seed"], len(case["counts"]), case["height"], case["width"], 4, 0, ragged=True,
                           counts=case["counts"])
    if case.get("tiny"):
        rng = np.random.RandomState(case["seed"] + 3)
        n = len(gt.boxes)
        cx, cy = rng.uniform(4, case["width"] - 4, n), rng.uniform(4, case["height"] - 4, n)
        w, h = rng.uniform(1.5, 4.0, n), rng.uniform(1.5, 4.0, n)
        b = np.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2], 1).astype(np.float32)
        gt = synth.GtBatch(b, gt.classes, gt.offsets)
    return gt


def geom_levels(g):
    return synth.level_sizes(g["height"], g["width"], g["bottom"], g["top"], g["mode"])


def case_gt(case) -> synth.GtBatch:
    g = GEOMETRIES[case["geom"]]
    counts = case.get("counts")
    gt = synth.gt_batch_np(case["seed"], case.get("batch", len(counts) if counts else 1), g["height"], g["width"],
                           case["classes"], case.get("max_gt", 0), ragged=True, counts=counts,
                           integer_coords=case.get("integer_coords", False))
    if case.get("flipped"):
        # unsanitised boxes like the reference's own test fixture (x2<x1 / y2<y1 allowed),
        # excluding w == h == 0 (NaN CIoU; outside the defined domain)
        rng = np.random.RandomState(case["seed"] + 7)
        b = rng.randint(0, g["height"], size=gt.boxes.shape).astype(np.float32)
        same = (b[:, 0] == b[:, 2]) & (b[:, 1] == b[:, 3])
        b[same, 2] += 3
        gt = synth.GtBatch(b, gt.classes, gt.offsets)
    return gt


def load(name):
    """Open ``tests/golden/<name>.npz`` as a dict of arrays."""
    with np.load(os.path.join(GOLDEN_DIR, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def train_maps(case):
    g = GEOMETRIES[case["geom"]]
    gt = case_gt(case)
    A = synth.num_anchors(geom_levels(g))
    return synth.dense_maps_np(case["seed"] + 1, gt.batch_size, A, case["classes"])


def forward_maps(case):
    g = GEOMETRIES[case["geom"]]
    A = synth.num_anchors(geom_levels(g))
    return synth.dense_maps_np(case["seed"], case["batch"], A, case["classes"], loc_mean=-2.0, loc_std=2.0)


def nms_inputs(case):
    return synth.nms_candidates_np(case["seed"], case["n"], case["size"], case["classes"])

"""numpy/ctypes front end of the CPU oracle (``oracle/od_oracle.c``).

TEST INFRASTRUCTURE ONLY — imported by ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs, never by
``sihl_b200``.  See the header of ``od_oracle.c`` for the parity pinning.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libod_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile ``od_oracle.c`` into ``oracle/_ref/`` (gcc, a second or two).  Content-stamped and file-locked so
    that copies with fresh file times and concurrent ranks do the right thing."""
    import fcntl
    import hashlib
    src = os.path.join(_HERE, "od_oracle.c")
    with open(src, "rb") as fh:
        fp = hashlib.sha256(fh.read()).hexdigest()
    stamp = os.path.join(_HERE, "_ref", "libod_oracle.stamp")

    def fresh():
        try:
            return os.path.exists(_LIB_PATH) and open(stamp).read().strip() == fp
        except OSError:
            return False

    if not force and fresh():
        return _LIB_PATH
    os.makedirs(os.path.join(_HERE, "_ref"), exist_ok=True)
    with open(os.path.join(_HERE, "_ref", ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if force or not fresh():
            subprocess.check_call(["make", "-C", _HERE, "-B", "-s"], stdout=subprocess.DEVNULL)
            with open(stamp, "w") as fh:
                fh.write(fp + "\n")
    return _LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_ciou_loss_row.restype = C.c_float
        _lib.orc_pos_loss.restype = C.c_int64
        _lib.orc_batched_nms.restype = C.c_int64
    return _lib


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _i64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int64)


def _i32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int32)


def anchors(levels: Sequence[Tuple[int, int]], img_w: int, img_h: int):
    """-> offsets [A,4], scales [A,4] (normalised), anchors [A,4] (px)."""
    hw = _i32(np.asarray(levels).reshape(-1, 2))
    A = int((hw[:, 0] * hw[:, 1]).sum())
    off, sc, an = (np.empty((A, 4), np.float32) for _ in range(3))
    rc = lib().orc_anchors(_p(hw), len(hw), int(img_w), int(img_h), _p(off), _p(sc), _p(an))
    assert rc == 0, rc
    return off, sc, an


def bbox_matching(anchors_px, gt, topk: int = 9, relative: bool = True):
    """Canonical ``bbox_matching`` -> (assignment i64 [A], iou f32 [A], best f32 [G])."""
    an, gt = _f32(anchors_px), _f32(gt).reshape(-1, 4)
    A, G = len(an), len(gt)
    assign, iou, best = np.empty(A, np.int64), np.empty(A, np.float32), np.zeros(G, np.float32)
    rc = lib().orc_bbox_matching(_p(an), C.c_int64(A), _p(gt), C.c_int64(G), int(topk), int(relative),
                                 _p(assign), _p(iou), _p(best))
    assert rc == 0, rc
    return assign, iou, best


def quad_matching(anchors_px, gt, topk: int = 9):
    """Canonical ``QuadrilateralDetection.bbox_matching`` -> (assignment i64, o2o bool, iou f32, rel f32), each [A]."""
    an, gt = _f32(anchors_px), _f32(gt).reshape(-1, 4)
    A, G = len(an), len(gt)
    assign, o2o = np.empty(A, np.int64), np.zeros(A, np.uint8)
    iou, rel = np.empty(A, np.float32), np.empty(A, np.float32)
    rc = lib().orc_quad_matching(_p(an), C.c_int64(A), _p(gt), C.c_int64(G), int(topk), _p(assign), _p(o2o), _p(iou), _p(rel))
    assert rc == 0, rc
    return assign, o2o.astype(bool), iou, rel


def assign_batch(anchors_px, gt_boxes, gt_offsets, topk: int = 9, relative: bool = True):
    an, gt, off = _f32(anchors_px), _f32(gt_boxes).reshape(-1, 4), _i32(gt_offsets)
    A, B = len(an), len(off) - 1
    assign, iou = np.empty((B, A), np.int64), np.empty((B, A), np.float32)
    best = np.zeros(len(gt), np.float32)
    rc = lib().orc_assign_batch(_p(an), C.c_int64(A), _p(gt), _p(off), B, int(topk), int(relative),
                                _p(assign), _p(iou), _p(best))
    assert rc == 0, rc
    return assign, iou, best


def dense_loss(loc, iou_pred, rel, sums: Optional[np.ndarray] = None) -> np.ndarray:
    loc, rel = _f32(loc).ravel(), _f32(rel).ravel()
    ip = None if iou_pred is None else _f32(iou_pred).ravel()
    sums = np.zeros(8, np.float64) if sums is None else sums
    lib().orc_dense_loss(_p(loc), _p(ip), _p(rel), C.c_int64(loc.size), _p(sums))
    return sums


def pos_loss(rel, assignment, offsets, scales, img_w, img_h, gt_boxes, gt_classes, gt_offsets,
             box_raw, cls_logits, dense_rows: bool, sums: Optional[np.ndarray] = None):
    """-> (sums, pos_index [P], box_loss_rows [P], ce_rows [P])."""
    rel, assignment = _f32(rel), _i64(assignment)
    B, A = rel.shape
    sums = np.zeros(8, np.float64) if sums is None else sums
    cap = int((rel > 0).sum())
    pos = np.empty(cap, np.int32)
    bl, ce = np.zeros(cap, np.float32), np.zeros(cap, np.float32)
    box_raw = None if box_raw is None else _f32(box_raw)
    cls_logits = None if cls_logits is None else _f32(cls_logits)
    ncls = 0 if cls_logits is None else cls_logits.shape[-1]
    P = lib().orc_pos_loss(_p(rel), _p(assignment), B, C.c_int64(A), _p(_f32(offsets)), _p(_f32(scales)),
                           int(img_w), int(img_h), _p(_f32(gt_boxes)), _p(_i64(gt_classes)), _p(_i32(gt_offsets)),
                           _p(box_raw), _p(cls_logits), ncls, int(dense_rows), _p(sums), _p(pos), _p(bl), _p(ce))
    assert P == cap
    return sums, pos, bl, ce


def loss_finalize(sums) -> np.ndarray:
    out = np.empty(5, np.float32)
    lib().orc_loss_finalize(_p(np.ascontiguousarray(sums, np.float64)), _p(out))
    return out


def ciou_loss_row(pred, tgt) -> float:
    return float(lib().orc_ciou_loss_row(_p(_f32(pred)), _p(_f32(tgt))))


def topk_rows(loc, K: int):
    loc = _f32(loc)
    B, A = loc.shape
    idx, top = np.empty((B, K), np.int64), np.empty((B, K), np.float32)
    rc = lib().orc_topk_rows(_p(loc), B, C.c_int64(A), int(K), _p(idx), _p(top))
    assert rc == 0, rc
    return idx, top


def decode_rows(top_logits, idx, cls_rows, box_rows, offsets, scales, img_w, img_h):
    top, idx = _f32(top_logits), _i64(idx)
    B, K = top.shape
    cls_rows, box_rows = _f32(cls_rows), _f32(box_rows)
    ncls = cls_rows.shape[-1]
    num = np.empty(B, np.int64)
    scores, classes, boxes = np.empty((B, K), np.float32), np.empty((B, K), np.int64), np.empty((B, K, 4), np.float32)
    lib().orc_decode_rows(_p(top), _p(idx), B, K, _p(cls_rows), ncls, _p(box_rows), _p(_f32(offsets)),
                          _p(_f32(scales)), int(img_w), int(img_h), _p(num), _p(scores), _p(classes), _p(boxes))
    return num, scores, classes, boxes


def batched_nms(boxes, scores, classes, iou_thr: float) -> np.ndarray:
    boxes, scores, classes = _f32(boxes).reshape(-1, 4), _f32(scores), _i64(classes)
    n = len(scores)
    keep = np.empty(n, np.int64)
    m = lib().orc_batched_nms(_p(boxes), _p(scores), _p(classes), C.c_int64(n), C.c_float(iou_thr), _p(keep))
    return keep[:m].copy()


def dense_postprocess(loc, cls, box_raw, offsets, scales, img_w, img_h,
                      score_thr: float = 0.05, iou_thr: float = 0.5, K: int = 100):
    loc, cls, box_raw = _f32(loc), _f32(cls), _f32(box_raw)
    B, A = loc.shape
    ncls = cls.shape[-1]
    num, ncand = np.empty(B, np.int64), np.empty(B, np.int64)
    scores, classes, boxes = np.empty((B, K), np.float32), np.empty((B, K), np.int64), np.empty((B, K, 4), np.float32)
    lib().orc_dense_postprocess(_p(loc), _p(cls), _p(box_raw), B, C.c_int64(A), ncls, _p(_f32(offsets)),
                                _p(_f32(scales)), int(img_w), int(img_h), C.c_float(score_thr), C.c_float(iou_thr),
                                int(K), _p(num), _p(scores), _p(classes), _p(boxes), _p(ncand))
    return num, scores, classes, boxes, ncand


def train_losses(anchors_px, offsets, scales, img_w, img_h, gt_boxes, gt_classes, gt_offsets,
                 loc, iou_pred, box_raw, cls_logits, dense_rows: bool = True, topk: int = 9):
    """Whole train path (a3-a10) -> dict(assignment, rel_iou, sums, losses, pos_index)."""
    assign, rel, best = assign_batch(anchors_px, gt_boxes, gt_offsets, topk, True)
    sums = dense_loss(loc, iou_pred, rel)
    sums, pos, bl, ce = pos_loss(rel, assign, offsets, scales, img_w, img_h, gt_boxes, gt_classes, gt_offsets,
                                 box_raw, cls_logits, dense_rows, sums)
    return dict(assignment=assign, rel_iou=rel, best_iou=best, sums=sums, losses=loss_finalize(sums),
                pos_index=pos, box_loss_rows=bl, ce_rows=ce)

"""od_math.h (the scalar arithmetic every kernel calls) compiled for the host and checked against
torch autograd / torchvision.  This is how the analytic backward is validated without a GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import torch
from torchvision import ops as tvops

from sihl_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    out = tmp_path_factory.mktemp("shim") / "libshim.so"
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off",
                           os.path.join(HERE, "host_math_shim.cpp"), "-o", str(out)])
    lib = C.CDLL(str(out))
    lib.shim_ciou_loss_row.restype = C.c_float
    lib.shim_bce_logits.restype = C.c_float
    lib.shim_bce_logits.argtypes = [C.c_float, C.c_float]
    return lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def test_ciou_matrix_matches_torchvision(shim):
    levels = synth.level_sizes(320, 320)
    from oracle import od_oracle as orc
    _, _, anchors = orc.anchors(levels, 320, 320)
    gt = synth.gt_boxes_np(np.random.RandomState(3), 40, 320, 320)
    out = np.empty((len(anchors), len(gt)), np.float32)
    shim.shim_ciou_matrix(_ptr(anchors), len(anchors), _ptr(gt), len(gt), _ptr(out))
    want = tvops.complete_box_iou(torch.from_numpy(anchors), torch.from_numpy(gt)).numpy()
    # same operator order; only atan differs in the last bit between glibc and torch-CPU (Sleef)
    np.testing.assert_allclose(out, want, rtol=0, atol=3e-7)
    assert (out == want).mean() > 0.9


def test_ciou_loss_value_and_grad_match_autograd(shim):
    rng = np.random.RandomState(11)
    n = 400
    c = rng.uniform(0.2, 0.8, (n, 2)); s = rng.uniform(0.02, 0.3, (n, 2))
    pred = np.concatenate([c - s, c + s], 1).astype(np.float32)
    c2 = c + rng.normal(0, 0.08, (n, 2)); s2 = s * np.exp(rng.normal(0, 0.4, (n, 2)))
    tgt = np.concatenate([c2 - s2, c2 + s2], 1).astype(np.float32)
    tgt[:20] += 5.0   # disjoint boxes: intersection branch off
    p = torch.tensor(pred, requires_grad=True)
    loss = tvops.complete_box_iou_loss(p, torch.from_numpy(tgt), reduction="none")
    loss.sum().backward()
    want_g = p.grad.numpy()
    got_l, got_g = np.empty(n, np.float32), np.empty((n, 4), np.float32)
    for i in range(n):
        g = np.empty(4, np.float32)
        got_l[i] = shim.shim_ciou_loss_row(_ptr(pred[i]), _ptr(tgt[i]), _ptr(g))
        got_g[i] = g
    np.testing.assert_allclose(got_l, loss.detach().numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(got_g, want_g, rtol=2e-4, atol=2e-4)


def test_bce_logits(shim):
    x = torch.linspace(-20, 20, 101)
    for t in (0.0, 1.0):
        want = torch.nn.functional.binary_cross_entropy_with_logits(x, torch.full_like(x, t), reduction="none")
        got = np.array([shim.shim_bce_logits(float(v), t) for v in x], np.float32)
        np.testing.assert_allclose(got, want.numpy(), rtol=1e-6, atol=1e-7)

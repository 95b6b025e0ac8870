"""The candidate windows of k_assign_select (csrc/od_assign.cu), emulated in numpy fp32 operation by operation, must be
SUPERSETS of the anchors with CIoU > 0 — checked by brute force against torchvision's complete_box_iou on the real anchor
tables (random, lattice-aligned, just-off-lattice, tiny and border ground truth; square / non-square / 3x3-P7
geometries).  The kernel itself is compared bit for bit with its all-pairs mode on the GPU (test_full_size_properties);
this test guards the window arithmetic where no GPU exists.  Keep `windows()` in sync with the kernel."""
import numpy as np
import pytest
import torch
from torchvision import ops as tvops
from oracle import torch_restatement as tr
from sihl_b200 import synth
f32=np.float32
def windows(gt, levels, W, H):
    x1,y1,x2,y2=[f32(v) for v in gt]
    area=f32(f32(x2-x1)*f32(y2-y1)); cx=f32(f32(x1+x2)/f32(2)); cy=f32(f32(y1+y2)/f32(2))
    out=[]
    for (lh,lw) in levels:
        isx=f32(f32(lw)/f32(W)); isy=f32(f32(lh)/f32(H)); sx=f32(f32(W)/f32(lw)); sy=f32(f32(H)/f32(lh))
        fw=f32(lw-1); fh=f32(lh-1)
        ex=f32(f32(0.02)+f32(1e-5)*max(abs(x1),abs(x2))); ey=f32(f32(0.02)+f32(1e-5)*max(abs(y1),abs(y2)))
        jlo=np.floor(f32(f32(x1-ex)*isx)); jhi=np.floor(f32(f32(x2+ex)*isx))
        ilo=np.floor(f32(f32(y1-ey)*isy)); ihi=np.floor(f32(f32(y2+ey)*isy))
        if area>0:
            cell=f32(sx*sy)
            iou_max=f32(min(cell,area)/max(cell,area))
            wg=f32(abs(f32(x2-x1))+sx); hg=f32(abs(f32(y2-y1))+sy)
            R=f32(np.sqrt(f32(iou_max*f32(f32(wg*wg)+f32(hg*hg)+f32(1))))*f32(1.02)+f32(0.01))
            jlo=max(jlo,np.floor(f32(f32(cx-R)*isx-f32(0.5)))); jhi=min(jhi,np.floor(f32(f32(cx+R)*isx-f32(0.5)))+1)
            ilo=max(ilo,np.floor(f32(f32(cy-R)*isy-f32(0.5)))); ihi=min(ihi,np.floor(f32(f32(cy+R)*isy-f32(0.5)))+1)
        j0=int(min(max(jlo,0),fw)); j1=int(min(max(jhi,0),fw)); i0=int(min(max(ilo,0),fh)); i1=int(min(max(ihi,0),fh))
        empty = (not jhi>=0) or (not jlo<=fw) or (not ihi>=0) or (not ilo<=fh) or j1<j0 or i1<i0
        out.append(None if empty else (i0,i1,j0,j1))
    return out


@pytest.mark.parametrize("H,W,mode", [(640, 640, "floor"), (320, 320, "ceil"), (384, 512, "floor"), (1024, 1024, "floor")])
def test_select_windows_contain_every_positive_ciou_anchor(H, W, mode):
    levels = synth.level_sizes(H, W, 3, 7, mode)
    anchors = tr.anchors_px(levels, W, H, "cpu")
    base = np.cumsum([0] + [h * w for h, w in levels])
    rng = np.random.RandomState(H)
    gts = list(synth.gt_batch_np(H, 10, H, W, 5, 40, ragged=False).boxes)
    for s in (8, 16, 32, 64, 128):
        for _ in range(16):
            a = rng.randint(0, W // s - 1) * s; b = rng.randint(0, H // s - 1) * s; k = rng.randint(1, 3)
            gts.append([a, b, min(W, a + k * s), min(H, b + k * s)])
            gts.append([a + 0.0001, b - 0.0001 if b > 0 else b, min(W, a + k * s) + 0.00005, min(H, b + k * s)])
    for _ in range(80):
        c = rng.uniform(0, W); d = rng.uniform(0, H)
        gts.append([c, d, min(W, c + rng.uniform(0.5, 6)), min(H, d + rng.uniform(0.5, 6))])
    gts = np.asarray(gts, np.float32)
    ciou = tvops.complete_box_iou(anchors, torch.from_numpy(gts)).clamp(0).numpy()
    checked = cand = 0
    for gi, gt in enumerate(gts):
        win = windows(gt, levels, W, H)
        for a in np.nonzero(ciou[:, gi] > 0)[0]:
            l = np.searchsorted(base, a, side="right") - 1
            i, j = divmod(a - base[l], levels[l][1])
            w_ = win[l]
            assert w_ is not None and w_[0] <= i <= w_[1] and w_[2] <= j <= w_[3], (gt, l, i, j, w_)
            checked += 1
        cand += sum(0 if w_ is None else (w_[1] - w_[0] + 1) * (w_[3] - w_[2] + 1) for w_ in win)
    assert checked > 3000
    assert cand / len(gts) < 90          # ~60 candidates per gt (the first version of the windows evaluated ~128)

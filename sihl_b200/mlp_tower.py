"""Row N4: run one of the head's ``torchvision.ops.MLP`` towers through the tcgen05 layer kernels.

The reference builds every tower as ``ops.MLP(num_channels, [num_channels] * num_layers + [out], norm_layer=nn.LayerNorm,
activation_layer=nn.SiLU)`` (ref src/sihl/heads/object_detection.py:51, :56-60), i.e. the module list
``Linear, LayerNorm, SiLU, Dropout`` x num_layers ``+ Linear, Dropout``.  :class:`PackedTower` reads those modules'
parameters once (bf16 weights, fp32 bias / gamma / beta, the last weight zero-padded to a tensor-core friendly row count)
and re-packs when any parameter's version counter moves; :func:`run_tower` chains ``ops.mlp_hidden`` x num_layers and
``ops.mlp_out`` over two ping-pong activation buffers (inference: no autograd graph is recorded).
:func:`run_tower_train` is the training path: the same kernels behind ``torch.autograd.Function`` nodes with a recomputing
backward, bf16 mixed precision (activations and operands bf16; accumulation, LayerNorm statistics and every parameter
gradient fp32) — the precision Lightning's ``precision="bf16-mixed"`` gives the torch towers.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
from torch import Tensor, nn

import os

from . import ops

# Training: keep every hidden layer's pre-activation (bf16 [M,256], written by the forward's epilogue) instead of
# recomputing it with one more GEMM pass in the backward: +512 B per location and layer of activation memory (2.2 GB at
# 640^2, batch 64, for the two dense towers) for one HBM round trip less per layer.  SIHL_MLP_SAVE_PRE=0 recomputes.
SAVE_PRE = os.environ.get("SIHL_MLP_SAVE_PRE", "1") != "0"


def _as(t: Tensor, dtype) -> Tensor:
    """``t.to(dtype)`` without the dispatcher round trip when there is nothing to convert (the usual case: fp32 parameters)."""
    return t if t.dtype == dtype else t.to(dtype)


def _f32(t: Tensor) -> Tensor:
    t = t.detach()
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


def _bf16(t: Tensor) -> Tensor:
    t = t.detach().to(torch.bfloat16)
    return t if t.is_contiguous() else t.contiguous()


class PackedTower:
    """bf16 / fp32 copies of one tower's parameters in the layout the kernels take."""

    def __init__(self, mlp: nn.Sequential) -> None:
        self.mlp = mlp
        self._versions: Optional[Tuple] = None
        self.hidden: List[Tuple[Tensor, Tensor, Tensor, Tensor, float]] = []
        self.out: Optional[Tuple[Tensor, Tensor, int]] = None

    @staticmethod
    def supported(mlp: nn.Module) -> bool:
        """True when ``mlp`` has the reference's tower structure at the width the kernels are built for."""
        try:
            layers = _split(mlp)
        except ValueError:
            return False
        hidden, last = layers
        ok = all(lin.in_features == ops.MLP_CHANNELS and lin.out_features == ops.MLP_CHANNELS and lin.bias is not None
                 and ln.elementwise_affine and tuple(ln.normalized_shape) == (ops.MLP_CHANNELS,) for lin, ln in hidden)
        return ok and last.in_features == ops.MLP_CHANNELS and last.out_features <= 256 and last.bias is not None

    def _current_versions(self) -> Tuple:
        return tuple((p.data_ptr(), p._version, p.device, p.dtype) for p in self.mlp.parameters())

    def refresh(self) -> "PackedTower":
        v = self._current_versions()
        if v == self._versions:
            return self
        hidden, last = _split(self.mlp)
        with torch.no_grad():
            self.hidden = [(lin.weight.detach().to(torch.bfloat16).contiguous(), lin.bias.detach().float().contiguous(),
                            ln.weight.detach().float().contiguous(), ln.bias.detach().float().contiguous(), float(ln.eps))
                           for lin, ln in hidden]
            n_out = last.out_features
            n_pad = ops.mlp_out_pad(n_out)
            w = torch.zeros((n_pad, ops.MLP_CHANNELS), dtype=torch.bfloat16, device=last.weight.device)
            w[:n_out] = last.weight.detach().to(torch.bfloat16)
            b = torch.zeros((n_pad,), dtype=torch.float32, device=last.weight.device)
            b[:n_out] = last.bias.detach().float()
            self.out = (w, b, n_out)
        self._versions = v
        return self


def _split(mlp: nn.Module):
    """(list of (Linear, LayerNorm) hidden pairs, last Linear) of a torchvision MLP; ValueError on any other structure."""
    mods = [m for m in mlp.children() if not isinstance(m, nn.Dropout)]
    hidden = []
    i = 0
    while i + 2 < len(mods):
        lin, ln, act = mods[i], mods[i + 1], mods[i + 2]
        if not (isinstance(lin, nn.Linear) and isinstance(ln, nn.LayerNorm) and isinstance(act, nn.SiLU)):
            raise ValueError("not a Linear -> LayerNorm -> SiLU tower")
        hidden.append((lin, ln))
        i += 3
    if i != len(mods) - 1 or not isinstance(mods[i], nn.Linear):
        raise ValueError("tower must end in one Linear")
    return hidden, mods[i]


def run_tower(packed: PackedTower, x: Tensor, scratch: Optional[Tuple[Tensor, Tensor]] = None) -> Tensor:
    """``mlp(x)`` for x [..., 256]: bf16 activations between layers, fp32 accumulation and normalisation, fp32 output
    [..., out_features].  ``scratch``: two bf16 [M,256] buffers to ping-pong between (allocated when omitted)."""
    packed.refresh()
    lead = x.shape[:-1]
    x2 = x.reshape(-1, x.shape[-1])
    cur = x2 if x2.dtype == torch.bfloat16 else x2.to(torch.bfloat16)
    cur = cur.contiguous()
    M = cur.shape[0]
    if scratch is None:
        scratch = (torch.empty((M, ops.MLP_CHANNELS), dtype=torch.bfloat16, device=x.device),
                   torch.empty((M, ops.MLP_CHANNELS), dtype=torch.bfloat16, device=x.device))
    for i, (w, b, g, be, eps) in enumerate(packed.hidden):
        cur = ops.mlp_hidden(cur, w, b, g, be, eps, out=scratch[i & 1][:M])
    w, b, n_out = packed.out
    y = ops.mlp_out(cur, w, b, n_out)
    return y.reshape(*lead, n_out)


# ---- training path: bf16 mixed precision (what Lightning's precision="bf16-mixed" gives the torch towers) ---------------
class _HiddenLayerFn(torch.autograd.Function):
    """``SiLU(LayerNorm(x W^T + b))`` with the forward on the tensor cores and a recomputing backward:
    v = x W^T + b again (tensor cores), LayerNorm + SiLU backward over rows (one HBM-bound kernel), dx = dv W (tensor
    cores), dW = dv^T x (a plain library GEMM)."""

    @staticmethod
    def forward(ctx, x, weight, bias, gamma, beta, eps):
        w16, b32, g32, be32 = _bf16(weight), _f32(bias), _f32(gamma), _f32(beta)
        if SAVE_PRE:
            y, stats, v = ops.mlp_hidden_train(x, w16, b32, g32, be32, eps, save_pre=True)
            ctx.save_for_backward(x, w16, b32, g32, be32, stats, v)
        else:
            y, stats = ops.mlp_hidden_train(x, w16, b32, g32, be32, eps)
            ctx.save_for_backward(x, w16, b32, g32, be32, stats)
        ctx.param_dtypes = (weight.dtype, bias.dtype, gamma.dtype, beta.dtype)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w16, b32, g32, be32, stats, *pre = ctx.saved_tensors
        dy = _as(dy, torch.bfloat16)
        dy = dy if dy.is_contiguous() else dy.contiguous()
        v = pre[0] if pre else ops.linear_bf16(x, w16, b32)                  # kept by the forward, or recomputed
        dv, dgamma, dbeta, dbias = ops.mlp_hidden_bwd(v, dy, stats, g32, be32)
        del v
        dx = ops.linear_bf16(dv, w16.t().contiguous(), ops.zero_bias(dv.device)) if ctx.needs_input_grad[0] else None   # dv W
        dw = torch.mm(dv.t(), x, out_dtype=torch.float32)                    # [out, in], fp32 out of the bf16 GEMM
        wd, bd, gd, bed = ctx.param_dtypes
        return dx, _as(dw, wd), _as(dbias, bd), _as(dgamma, gd), _as(dbeta, bed), None


class _OutLayerFn(torch.autograd.Function):
    """The tower's last Linear: forward on the tensor cores (fp32 result), backward as bf16 matrix products."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        n_out = weight.shape[0]
        n_pad = ops.mlp_out_pad(n_out)
        w = torch.zeros((n_pad, ops.MLP_CHANNELS), dtype=torch.bfloat16, device=x.device)
        w[:n_out] = weight.detach().to(torch.bfloat16)
        b = torch.zeros((n_pad,), dtype=torch.float32, device=x.device)
        b[:n_out] = bias.detach().float()
        ctx.save_for_backward(x, w[:n_out])
        ctx.param_dtypes = (weight.dtype, bias.dtype)
        return ops.mlp_out(x, w, b, n_out)

    @staticmethod
    def backward(ctx, dout):
        x, w16 = ctx.saved_tensors
        d16 = dout.to(torch.bfloat16)
        dx = torch.matmul(d16, w16) if ctx.needs_input_grad[0] else None      # [M, 256] bf16
        dw = torch.mm(d16.t(), x, out_dtype=torch.float32)
        wd, bd = ctx.param_dtypes
        return dx, dw.to(wd), dout.float().sum(0).to(bd)


class _LastHiddenOutFn(torch.autograd.Function):
    """A tower's last hidden layer AND its single-output Linear (the location and IoU towers, ref :56, :60) as one node: the
    backward never materialises the [M,256] gradient between the two — ``d16[:, None] * w_out`` would come out of a K = 1
    library GEMM only to be read back by the LayerNorm + SiLU backward, which forms it in registers instead
    (``ops.mlp_hidden_bwd_rank1``).  Same arithmetic and roundings as ``_OutLayerFn`` after ``_HiddenLayerFn``."""

    @staticmethod
    def forward(ctx, x, weight, bias, gamma, beta, eps, w_out, b_out):
        w16, b32, g32, be32 = _bf16(weight), _f32(bias), _f32(gamma), _f32(beta)
        if SAVE_PRE:
            y, stats, v = ops.mlp_hidden_train(x, w16, b32, g32, be32, eps, save_pre=True)
        else:
            (y, stats), v = ops.mlp_hidden_train(x, w16, b32, g32, be32, eps), None
        n_pad = ops.mlp_out_pad(1)
        wo = torch.zeros((n_pad, ops.MLP_CHANNELS), dtype=torch.bfloat16, device=x.device)
        wo[:1] = w_out.detach().to(torch.bfloat16)
        bo = torch.zeros((n_pad,), dtype=torch.float32, device=x.device)
        bo[:1] = b_out.detach().float()
        ctx.save_for_backward(x, w16, b32, g32, be32, stats, y, wo, *([v] if v is not None else []))
        ctx.param_dtypes = (weight.dtype, bias.dtype, gamma.dtype, beta.dtype, w_out.dtype, b_out.dtype)
        return ops.mlp_out(y, wo, bo, 1)

    @staticmethod
    def backward(ctx, dout):
        x, w16, b32, g32, be32, stats, y, wo, *pre = ctx.saved_tensors
        dcol = dout.reshape(-1).float().contiguous()                          # [M] fp32
        d16 = dcol.to(torch.bfloat16)
        v = pre[0] if pre else ops.linear_bf16(x, w16, b32)                  # kept by the forward, or recomputed
        dv, dgamma, dbeta, dbias = ops.mlp_hidden_bwd_rank1(v, dcol, wo[0], stats, g32, be32)
        del v
        dx = ops.linear_bf16(dv, w16.t().contiguous(), ops.zero_bias(dv.device)) if ctx.needs_input_grad[0] else None
        dw = torch.mm(dv.t(), x, out_dtype=torch.float32)
        dw_out = torch.mm(d16[None, :], y, out_dtype=torch.float32)          # [1, 256]
        wd, bd, gd, bed, wod, bod = ctx.param_dtypes
        return dx, _as(dw, wd), _as(dbias, bd), _as(dgamma, gd), _as(dbeta, bed), None, _as(dw_out, wod), _as(dcol.sum(0, keepdim=True), bod)


class _FanOutFn(torch.autograd.Function):
    """The towers' common input handed to its three consumers — the location tower, the IoU tower (both dense) and the
    gather of the positives' rows for the box / class towers (ref :157, :175, :184) — as ONE autograd node.  Left to
    itself autograd materialises zeros [B*A, 256] for the gather's backward, index-adds into them and then sums three
    full-size gradients pairwise (fill + index_add + add + add_: 0.41 ms at cfg1); here the backward is one dense add
    and one in-place index_add of the few positive rows.

    ``apply(flat2d [B*A, C], pos_index i32 [cap])`` -> (flat2d, flat2d, flat2d[pos_index])."""

    @staticmethod
    def forward(ctx, flat2d, pos_index):
        ctx.save_for_backward(pos_index)
        ctx.shape, ctx.dtype = flat2d.shape, flat2d.dtype
        return flat2d.view_as(flat2d), flat2d.view_as(flat2d), flat2d.index_select(0, pos_index)

    @staticmethod
    def backward(ctx, d_a, d_b, d_rows):
        (pos_index,) = ctx.saved_tensors
        if d_a is not None and d_b is not None:
            g = d_a + d_b
        elif d_a is not None or d_b is not None:
            g = (d_a if d_a is not None else d_b).clone()
        else:
            g = torch.zeros(ctx.shape, dtype=ctx.dtype, device=pos_index.device)
        if d_rows is not None:
            g.index_add_(0, pos_index, d_rows.to(g.dtype))       # padding rows repeat row 0 with a zero gradient
        return g, None


class _ToBf16Fn(torch.autograd.Function):
    """fp32 -> bf16 at the tower's entrance; the backward hands the gradient back in fp32 (one full-width kernel)."""

    @staticmethod
    def forward(ctx, x):
        ctx.src_dtype = x.dtype
        return x.to(torch.bfloat16)

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        return ops.bf16_to_f32(g) if ctx.src_dtype == torch.float32 and g.dtype == torch.bfloat16 else g.to(ctx.src_dtype)


def run_tower_train(mlp: nn.Sequential, x: Tensor) -> Tensor:
    """``mlp(x)`` with gradients, bf16 activations / operands and fp32 accumulation, statistics and parameters' gradients."""
    hidden, last = _split(mlp)
    lead = x.shape[:-1]
    cur = x.reshape(-1, x.shape[-1]).contiguous()
    if cur.dtype != torch.bfloat16:
        cur = _ToBf16Fn.apply(cur)
    fuse_last = last.out_features == 1 and len(hidden) > 0
    for lin, ln in (hidden[:-1] if fuse_last else hidden):
        cur = _HiddenLayerFn.apply(cur, lin.weight, lin.bias, ln.weight, ln.bias, float(ln.eps))
    if fuse_last:
        lin, ln = hidden[-1]
        y = _LastHiddenOutFn.apply(cur, lin.weight, lin.bias, ln.weight, ln.bias, float(ln.eps), last.weight, last.bias)
    else:
        y = _OutLayerFn.apply(cur, last.weight, last.bias)
    return y.reshape(*lead, last.out_features)


class _LateralsTrainFn(torch.autograd.Function):
    """All laterals in training mode — per level ``Conv2d(256, 256, 1, bias=False)`` + batch-statistics ``BatchNorm2d`` (ref
    object_detection.py:52-55, :102-105) — each written straight into its slice of the concatenated bf16 features.

    Forward: the conv output is never materialised.  Its batch mean and variance follow from the input's first and second
    moments (mean = W s / M, E[y^2] = diag(W G W^T) / M with s = sum of the rows, G = rows^T rows, fp32), so conv +
    BatchNorm is one GEMM pass with folded weights, like the eval path.  Backward: the normalised conv output is recomputed
    by the same GEMM, the BatchNorm backward runs over rows in two HBM-bound kernels, the conv's dgrad is the tensor-core
    GEMM (+ the layout change back to NCHW fp32) and its wgrad a library GEMM.  Also returns every level's (mean, biased
    variance) so the caller can update the modules' running statistics exactly as torch does.

    ``apply(eps_per_level, x_0, w_0, gamma_0, beta_0, x_1, ...)`` -> (flat [B, A, 256] bf16, mean_0, var_0, mean_1, ...)."""

    @staticmethod
    def forward(ctx, eps, *tensors):
        n_levels = len(tensors) // 4
        xs, ws, gs, bs = tensors[0::4], tensors[1::4], tensors[2::4], tensors[3::4]
        B = xs[0].shape[0]
        sizes = [(int(x.shape[2]), int(x.shape[3])) for x in xs]
        flat = torch.empty((B, sum(h * w for h, w in sizes), ops.MLP_CHANNELS), dtype=torch.bfloat16, device=xs[0].device)
        saved, stats, offset = [], [], 0
        for lvl in range(n_levels):
            h, w = sizes[lvl]
            M = B * h * w
            rows = ops.lateral_rows(xs[lvl].detach().contiguous())
            w16 = ws[lvl].detach()[:, :, 0, 0].to(torch.bfloat16).contiguous()
            wf = w16.float()
            s = ops.rows_colsum(rows)
            gram = torch.mm(rows.t(), rows, out_dtype=torch.float32)
            mean = (wf @ s) / M
            var = ((((wf @ gram) * wf).sum(1) / M) - mean * mean).clamp_min(0.0)
            invstd = torch.rsqrt(var + eps[lvl])
            scale = gs[lvl].detach().float() * invstd
            ops.lateral_linear(rows, (wf * scale[:, None]).to(torch.bfloat16).contiguous(),
                               (bs[lvl].detach().float() - mean * scale).contiguous(), h * w, flat, offset)
            saved += [rows, w16, mean, invstd, scale]
            stats += [mean, var]
            offset += h * w
        ctx.save_for_backward(*saved)
        ctx.sizes, ctx.batch = sizes, B
        ctx.dtypes = [(ws[l].dtype, gs[l].dtype, bs[l].dtype) for l in range(n_levels)]
        ctx.mark_non_differentiable(*stats)
        return (flat, *stats)

    @staticmethod
    def backward(ctx, dflat, *_unused):
        saved = ctx.saved_tensors
        B, grads, offset = ctx.batch, [None], 0
        dflat = dflat.to(torch.bfloat16).contiguous()                         # [B, A, 256]: every level reads its slice in place
        for lvl, (h, w) in enumerate(ctx.sizes):
            rows, w16, mean, invstd, scale = saved[5 * lvl:5 * lvl + 5]
            wf = w16.float()
            n = ops.linear_bf16(rows, (wf * invstd[:, None]).to(torch.bfloat16).contiguous(), (-mean * invstd).contiguous())
            dy, dgamma, dbeta = ops.bn_bwd_rows(dflat, n, scale, dz_rows_per_image=int(dflat.shape[1]), dz_row_offset=offset,
                                                rows_per_image=h * w)
            del n
            dx = None
            if ctx.needs_input_grad[1 + 4 * lvl]:
                dx = ops.rows_to_nchw(ops.linear_bf16(dy, w16.t().contiguous(), ops.zero_bias(dy.device)), B, h, w)
            dw = torch.mm(dy.t(), rows, out_dtype=torch.float32)[:, :, None, None]
            wd, gd, bd = ctx.dtypes[lvl]
            grads += [dx, dw.to(wd), dgamma.to(gd), dbeta.to(bd)]
            offset += h * w
        return tuple(grads)

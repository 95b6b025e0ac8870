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

from . import ops


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
        w16 = weight.detach().to(torch.bfloat16).contiguous()
        b32, g32, be32 = bias.detach().float().contiguous(), gamma.detach().float().contiguous(), beta.detach().float().contiguous()
        y, stats = ops.mlp_hidden_train(x, w16, b32, g32, be32, eps)
        ctx.save_for_backward(x, w16, b32, g32, be32, stats)
        ctx.param_dtypes = (weight.dtype, bias.dtype, gamma.dtype, beta.dtype)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w16, b32, g32, be32, stats = ctx.saved_tensors
        dy = dy.to(torch.bfloat16).contiguous()
        v = ops.linear_bf16(x, w16, b32)                                     # recompute the pre-activation
        dv, dgamma, dbeta, dbias = ops.mlp_hidden_bwd(v, dy, stats, g32, be32)
        del v
        dx = ops.linear_bf16(dv, w16.t().contiguous(), torch.zeros_like(b32)) if ctx.needs_input_grad[0] else None   # dv W
        dw = torch.matmul(dv.t(), x).float()                                 # [out, in]
        wd, bd, gd, bed = ctx.param_dtypes
        return dx, dw.to(wd), dbias.to(bd), dgamma.to(gd), dbeta.to(bed), None


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
        dw = torch.matmul(d16.t(), x).float()
        wd, bd = ctx.param_dtypes
        return dx, dw.to(wd), dout.float().sum(0).to(bd)


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
    for lin, ln in hidden:
        cur = _HiddenLayerFn.apply(cur, lin.weight, lin.bias, ln.weight, ln.bias, float(ln.eps))
    y = _OutLayerFn.apply(cur, last.weight, last.bias)
    return y.reshape(*lead, last.out_features)

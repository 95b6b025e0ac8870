"""Row N4: the head's per-location MLP towers on the tensor cores (sihl_od_mlp_hidden / sihl_od_mlp_out)
against torch.  The reference builds the towers at src/sihl/heads/object_detection.py:51-61 and applies them at :116,
:121 and :175.  Tolerances: the output layer accumulates bf16 products in fp32 exactly like an fp64 reference on the
same bf16 operands (<= 2e-5 abs); a hidden layer's result is rounded to bf16 once (half an ulp: 2^-9 relative, plus
tanh.approx's 2^-11 in SiLU)."""
import pytest
import torch
import torch.nn.functional as F
from torch import nn
from torchvision import ops as tvops

from sihl_b200 import ops
from sihl_b200.mlp_tower import PackedTower, run_tower

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rand(shape, seed, scale=1.0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    return torch.randn(shape, generator=g, device=DEV) * scale


def _hidden_ref(x, w, b, gamma, beta, eps=1e-5):
    pre = (x.double() @ w.double().T + b.double()).float()
    return F.silu(F.layer_norm(pre, (256,), gamma, beta, eps))


@pytest.mark.parametrize("M", [1, 127, 128, 129, 300, 128 * 148 + 5, 128 * 400])
def test_hidden_layer(M):
    """Single tiles, ragged tails (rows past M are zero-filled by the TMA unit and never stored), one tile per SM plus
    a tail, and three tiles per CTA (both TMEM stages, ring wrap-around)."""
    x = _rand((M, 256), 1).bfloat16()
    w = _rand((256, 256), 2, 1 / 16).bfloat16()
    b, gamma, beta = _rand((256,), 3), 1 + 0.1 * _rand((256,), 4), 0.1 * _rand((256,), 5)
    guard = torch.full((M + 2, 256), 7.0, dtype=torch.bfloat16, device=DEV)          # rows M, M+1 must stay untouched
    y = ops.mlp_hidden(x, w, b, gamma, beta, out=guard[:M])
    ref = _hidden_ref(x, w, b, gamma, beta)
    torch.testing.assert_close(y.float(), ref, rtol=2 ** -8, atol=2e-3)
    assert (guard[M:] == 7.0).all()


@pytest.mark.parametrize("out_features", [1, 4, 16, 80, 200])
@pytest.mark.parametrize("M", [1, 300, 128 * 300 + 17])
def test_output_layer(out_features, M):
    x = _rand((M, 256), 6).bfloat16()
    n_pad = ops.mlp_out_pad(out_features)
    w = torch.zeros((n_pad, 256), dtype=torch.bfloat16, device=DEV)
    w[:out_features] = _rand((out_features, 256), 7, 1 / 16).bfloat16()
    b = torch.zeros((n_pad,), device=DEV)
    b[:out_features] = _rand((out_features,), 8)
    guard = torch.full((M + 1, out_features), -3.0, device=DEV)
    y = ops.mlp_out(x, w, b, out_features, out=guard[:M])
    ref = (x.double() @ w[:out_features].double().T + b[:out_features].double()).float()
    torch.testing.assert_close(y, ref, rtol=0, atol=2e-5)
    assert (guard[M:] == -3.0).all()


def test_extreme_rows():
    """Constant rows (variance 0 -> rstd = 1/sqrt(eps), as torch), large magnitudes, and a row of zeros."""
    x = torch.zeros((4, 256), device=DEV)
    x[1] = 1.0
    x[2] = _rand((256,), 9) * 100
    x[3, ::2] = 50.0
    x = x.bfloat16()
    w = _rand((256, 256), 10, 1 / 16).bfloat16()
    b, gamma, beta = torch.zeros(256, device=DEV), torch.ones(256, device=DEV), torch.zeros(256, device=DEV)
    y = ops.mlp_hidden(x, w, b, gamma, beta)
    ref = _hidden_ref(x, w, b, gamma, beta)
    assert torch.isfinite(y).all()
    torch.testing.assert_close(y.float(), ref, rtol=2 ** -8, atol=2e-3)


@pytest.mark.parametrize("out_features", [1, 4, 80])
def test_tower_matches_the_torchvision_module(out_features):
    """The whole tower (4 hidden layers + output, as the reference builds it) against the torch module: tight against the
    same module run with bf16 rounding between layers, loose (bf16 chain vs fp32 chain) against the fp32 module."""
    torch.manual_seed(0)
    mlp = tvops.MLP(256, [256] * 4 + [out_features], norm_layer=nn.LayerNorm, activation_layer=nn.SiLU).to(DEV).eval()
    with torch.no_grad():
        for m in mlp:
            if isinstance(m, nn.LayerNorm):
                m.weight.add_(0.1 * torch.randn_like(m.weight)); m.bias.add_(0.1 * torch.randn_like(m.bias))
    assert PackedTower.supported(mlp)
    x = _rand((3, 1000, 256), 11)
    packed = PackedTower(mlp)
    y = run_tower(packed, x)
    assert y.shape == (3, 1000, out_features) and y.dtype == torch.float32
    with torch.no_grad():
        cur = x.bfloat16().float()
        mods = [m for m in mlp if not isinstance(m, nn.Dropout)]
        for i in range(0, len(mods) - 1, 3):
            lin, ln = mods[i], mods[i + 1]
            pre = (cur.double() @ lin.weight.bfloat16().double().T + lin.bias.double()).float()
            cur = F.silu(ln(pre)).bfloat16().float()
        emu = (cur.double() @ mods[-1].weight.bfloat16().double().T + mods[-1].bias.double()).float()
        full = mlp(x)
    torch.testing.assert_close(y, emu, rtol=2e-2, atol=2e-2)           # bf16 roundings may flip an ulp along the chain
    assert (y - full).abs().max() < 0.1 and (y - full).abs().mean() < 0.01
    # re-packing follows an optimizer step
    with torch.no_grad():
        mlp[0].weight.mul_(0.5)
    y2 = run_tower(packed, x)
    assert (y2 - y).abs().max() > 1e-3
    with torch.no_grad():
        torch.testing.assert_close(y2, mlp(x), rtol=0, atol=0.1)


def test_rejects_what_the_kernels_are_not_built_for():
    x = torch.zeros((8, 128), dtype=torch.bfloat16, device=DEV)
    w = torch.zeros((128, 128), dtype=torch.bfloat16, device=DEV)
    v = torch.zeros((128,), device=DEV)
    with pytest.raises(ValueError):
        ops.mlp_hidden(x, w, v, v, v)
    with pytest.raises(ValueError):
        ops.mlp_out_pad(300)
    mlp = tvops.MLP(128, [128, 4], norm_layer=nn.LayerNorm, activation_layer=nn.SiLU)
    assert not PackedTower.supported(mlp)
    assert not PackedTower.supported(tvops.MLP(256, [256, 4], norm_layer=None, activation_layer=nn.ReLU))


def _head(num_classes=12):
    from sihl_b200.heads import ObjectDetection
    torch.manual_seed(0)
    return ObjectDetection(in_channels=[3] + [64] * 5, num_classes=num_classes, bottom_level=3, top_level=5, num_channels=256,
                           num_layers=4, max_instances=50).to(DEV).eval()


def _pyramid(batch=2, size=256):
    g = torch.Generator().manual_seed(3)
    return [torch.randn((batch, 3, size, size), generator=g).to(DEV)] + [
        torch.randn((batch, 64, size // 2 ** l, size // 2 ** l), generator=g).to(DEV) for l in range(1, 6)]


def test_head_inference_through_the_tensor_core_towers():
    """``mlp_backend = "tcgen05"``: forward / postprocess / get_saliency under no_grad run the towers through the kernels;
    results agree with the torch towers to bf16-chain accuracy, and a training step (gradients recorded) is untouched."""
    model, inputs = _head(), _pyramid()
    with torch.no_grad():
        model.loc_head[-2].bias.fill_(-1.0)                    # scores around 0.3: candidates exist for postprocess
        ref_num, ref_scores, ref_cls, ref_boxes = model.forward(inputs)
        ref_sal = model.get_saliency(inputs)
        ref_post = model.postprocess(inputs, 0.05, 0.5)
        model.mlp_backend = "tcgen05"
        num, scores, cls, boxes = model.forward(inputs)
        sal = model.get_saliency(inputs)
        post = model.postprocess(inputs, 0.05, 0.5)
    assert set(model._packed_towers) == {"loc_head", "cls_head", "box_head"}          # the kernels did run
    assert scores.shape == ref_scores.shape and boxes.shape == ref_boxes.shape and cls.dtype == torch.int64
    torch.testing.assert_close(scores, ref_scores, rtol=0, atol=0.02)                 # sorted top-K scores
    torch.testing.assert_close(sal, ref_sal, rtol=0, atol=0.02)
    assert (num - ref_num).abs().max() <= 2
    torch.testing.assert_close(post[1][:, :5], ref_post[1][:, :5], rtol=0, atol=0.02)  # best detections after NMS
    # gradients recorded -> torch modules, identical to the default backend
    model.train()
    tgt = {"classes": [torch.tensor([1, 2], device=DEV), torch.tensor([3], device=DEV)],
           "boxes": [torch.tensor([[10., 20., 100., 120.], [50., 60., 200., 220.]], device=DEV), torch.tensor([[30., 30., 90., 90.]], device=DEV)]}
    loss_a, _ = model.training_step(inputs, **tgt)
    model.mlp_backend = "torch"
    loss_b, _ = model.training_step(inputs, **tgt)
    assert loss_a.item() == loss_b.item() and loss_a.requires_grad


def test_head_rejects_the_backend_at_other_widths():
    from sihl_b200.heads import ObjectDetection
    model = ObjectDetection(in_channels=[3] + [32] * 5, num_classes=4, num_channels=64, num_layers=2).to(DEV).eval()
    model.mlp_backend = "tcgen05"
    g = torch.Generator().manual_seed(3)
    inputs = [torch.randn((1, 3, 64, 64), generator=g).to(DEV)] + [torch.randn((1, 32, 64 // 2 ** l, 64 // 2 ** l), generator=g).to(DEV) for l in range(1, 6)]
    with torch.no_grad(), pytest.raises(ValueError):
        model.forward(inputs)


@pytest.mark.parametrize("hw", [(80, 80), (5, 5), (13, 7)])
def test_lateral_rows_is_the_rearrange_of_the_reference(hw):
    """ref :105 ``rearrange(x, "b c h w -> b (h w) c")`` + bf16 rounding, for sizes that are and are not multiples of 64."""
    h, w = hw
    x = _rand((3, 256, h, w), 21)
    rows = ops.lateral_rows(x)
    want = x.flatten(2).transpose(1, 2).reshape(-1, 256).bfloat16()
    assert torch.equal(rows, want)


def test_lateral_linear_lands_in_its_slice_of_the_concatenated_features():
    B, A = 3, 700
    out = torch.full((B, A, 256), 9.0, dtype=torch.bfloat16, device=DEV)
    wt = _rand((256, 256), 22, 1 / 16).bfloat16()
    bias = _rand((256,), 23)
    for hw_l, off in ((400, 0), (100, 400), (150, 500)):                     # 650 of the 700 rows per image get written
        rows = _rand((B * hw_l, 256), 24 + off).bfloat16()
        ops.lateral_linear(rows, wt, bias, hw_l, out, off)
        ref = (rows.double() @ wt.double().T + bias.double()).float().reshape(B, hw_l, 256)
        torch.testing.assert_close(out[:, off:off + hw_l].float(), ref, rtol=2 ** -8, atol=2e-3)
    assert (out[:, 650:] == 9.0).all()


def test_head_laterals_on_the_tensor_cores_match_conv_plus_batchnorm():
    """Eval-mode laterals (1x1 conv + BatchNorm with non-trivial running statistics) through lateral_rows + lateral_linear
    against the torch modules; in training mode (batch statistics) the path is not taken."""
    from sihl_b200.heads import ObjectDetection
    torch.manual_seed(1)
    model = ObjectDetection(in_channels=[3, 16, 32, 256, 256, 256], num_classes=5, num_channels=256, num_layers=1).to(DEV)
    with torch.no_grad():
        for lat in model.laterals:
            lat[1].running_mean.normal_(0, 0.5); lat[1].running_var.uniform_(0.5, 2.0)
            lat[1].weight.normal_(1, 0.2); lat[1].bias.normal_(0, 0.2)
    model.eval()
    g = torch.Generator().manual_seed(5)
    inputs = [torch.randn((2, c, max(1, 96 // 2 ** l), max(1, 96 // 2 ** l)), generator=g).to(DEV) for l, c in enumerate(model.in_channels)]
    model.mlp_backend = "tcgen05"
    with torch.no_grad():
        flat = model._tower_feats(inputs)
        ref = model._flat_feats(inputs)
    assert flat.dtype == torch.bfloat16 and flat.shape == ref.shape
    torch.testing.assert_close(flat.float(), ref, rtol=2e-2, atol=2e-2)       # bf16 operands and result
    model.train()
    with torch.no_grad():
        assert model._folded_laterals() is None


@pytest.mark.parametrize("backend", ["torch", "tcgen05"])
def test_graphed_forward_replays_the_same_detections(backend):
    """sihl_b200.serving.GraphedInference: forward captured once, replayed on new inputs — same outputs as the eager call."""
    from sihl_b200.serving import GraphedInference
    model = _head()
    model.mlp_backend = backend
    a, b = _pyramid(), [t * 0.5 + 0.1 for t in _pyramid()]
    graphed = GraphedInference(model.forward, a)
    for inputs in (a, b, a):
        with torch.no_grad():
            want = [t.clone() for t in model.forward(inputs)]
        got = graphed(inputs)
        for g, w in zip(got, want):
            assert torch.equal(g, w)
    with pytest.raises(ValueError):
        graphed([t[:1] for t in a])


def _rel_l2(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-12))


def _cos(a, b):
    return float(F.cosine_similarity(a.float().flatten(), b.float().flatten(), dim=0))


@pytest.mark.parametrize("M", [300, 128 * 150 + 7])
def test_hidden_layer_backward_against_torch_autograd(M):
    """Training path of one hidden layer (forward on the tensor cores + recomputing backward) against torch fp32 autograd on
    the same bf16-rounded operands.  bf16 mixed precision: gradients agree in direction (cosine) and to a few 1e-3 in
    relative L2 norm; statistics and parameter gradients are accumulated in fp32."""
    from sihl_b200.mlp_tower import _HiddenLayerFn
    lin, ln = nn.Linear(256, 256).to(DEV), nn.LayerNorm(256).to(DEV)
    with torch.no_grad():
        ln.weight.add_(0.2 * torch.randn_like(ln.weight)); ln.bias.add_(0.2 * torch.randn_like(ln.bias))
    x = _rand((M, 256), 31).bfloat16().requires_grad_(True)
    gout = _rand((M, 256), 32).bfloat16()
    y = _HiddenLayerFn.apply(x, lin.weight, lin.bias, ln.weight, ln.bias, ln.eps)
    y.backward(gout)
    ours = [x.grad.clone(), lin.weight.grad.clone(), lin.bias.grad.clone(), ln.weight.grad.clone(), ln.bias.grad.clone()]
    for p in (lin.weight, lin.bias, ln.weight, ln.bias):
        p.grad = None
    xr = x.detach().float().requires_grad_(True)
    wr = lin.weight.detach().bfloat16().float().requires_grad_(True)
    yr = F.silu(F.layer_norm(F.linear(xr, wr, lin.bias), (256,), ln.weight, ln.bias, ln.eps))
    torch.testing.assert_close(y.float(), yr, rtol=2 ** -8, atol=2e-3)
    yr.backward(gout.float())
    refs = [xr.grad, wr.grad, lin.bias.grad, ln.weight.grad, ln.bias.grad]
    for name, a, b in zip(("dx", "dW", "dbias", "dgamma", "dbeta"), ours, refs):
        assert _cos(a, b) > 0.9995, (name, _cos(a, b))
        assert _rel_l2(a, b) < 2e-2, (name, _rel_l2(a, b))


def test_tower_training_path_against_the_torch_module():
    """run_tower_train (4 hidden layers + output) against the torchvision module in fp32: loss and every parameter gradient."""
    from sihl_b200.mlp_tower import run_tower_train
    torch.manual_seed(0)
    mlp = tvops.MLP(256, [256] * 4 + [4], norm_layer=nn.LayerNorm, activation_layer=nn.SiLU).to(DEV)
    x = _rand((2, 3000, 256), 41).requires_grad_(True)
    tgt = _rand((2, 3000, 4), 42)
    loss = F.mse_loss(run_tower_train(mlp, x), tgt)
    loss.backward()
    ours = {n: p.grad.clone() for n, p in mlp.named_parameters()}
    dx = x.grad.clone()
    mlp.zero_grad(); x.grad = None
    ref_loss = F.mse_loss(mlp(x), tgt)
    ref_loss.backward()
    assert loss.item() == pytest.approx(ref_loss.item(), rel=2e-2)
    assert _cos(dx, x.grad) > 0.995 and _rel_l2(dx, x.grad) < 0.1
    for n, p in mlp.named_parameters():
        assert _cos(ours[n], p.grad) > 0.995, (n, _cos(ours[n], p.grad))
        assert _rel_l2(ours[n], p.grad) < 0.1, (n, _rel_l2(ours[n], p.grad))


@pytest.mark.parametrize("M", [1, 129, 128 * 150 + 7])
def test_training_forward_keeps_the_pre_activation(M):
    """``ops.mlp_hidden_train(save_pre=True)``: y and the row statistics are those of the plain training forward, and the
    stored pre-activation is bit-equal to what the backward used to recompute (``ops.linear_bf16``); rows past M of a
    guard-banded buffer stay untouched."""
    x = _rand((M, 256), 91).bfloat16()
    w = _rand((256, 256), 92, 1 / 16).bfloat16()
    b, gamma, beta = _rand((256,), 93), 1 + 0.1 * _rand((256,), 94), 0.1 * _rand((256,), 95)
    y0, st0 = ops.mlp_hidden_train(x, w, b, gamma, beta)
    y1, st1, v = ops.mlp_hidden_train(x, w, b, gamma, beta, save_pre=True)
    assert torch.equal(y0, y1) and torch.equal(st0, st1)
    assert torch.equal(v, ops.linear_bf16(x, w, b))


@pytest.mark.parametrize("out_features", [1, 4])
def test_kept_pre_activation_gives_the_gradients_of_the_recompute(out_features, monkeypatch):
    """``mlp_tower.SAVE_PRE`` (pre-activations written by the forward) against the recomputing backward: the stored v is
    what the recompute produces, so every gradient is bit-equal."""
    from sihl_b200 import mlp_tower
    torch.manual_seed(2)
    mlp = tvops.MLP(256, [256] * 2 + [out_features], norm_layer=nn.LayerNorm, activation_layer=nn.SiLU).to(DEV)
    x0 = _rand((1500, 256), 97)
    tgt = _rand((1500, out_features), 98)
    grads = []
    for keep in (True, False):
        monkeypatch.setattr(mlp_tower, "SAVE_PRE", keep)
        mlp.zero_grad()
        x = x0.clone().requires_grad_(True)
        F.mse_loss(mlp_tower.run_tower_train(mlp, x), tgt).backward()
        grads.append([x.grad.clone()] + [p.grad.clone() for p in mlp.parameters()])
    for a, b in zip(*grads):
        assert torch.equal(a, b)


@pytest.mark.parametrize("M", [1, 300, 128 * 150 + 7])
def test_rank1_backward_equals_the_materialised_outer_product(M):
    """``ops.mlp_hidden_bwd_rank1`` (upstream gradient bf16(dout) x w_out formed in registers) == ``ops.mlp_hidden_bwd`` on
    the [M,256] matrix the K = 1 library GEMM produces: same roundings, same accumulation order -> bit-equal."""
    v = _rand((M, 256), 51).bfloat16()
    x = _rand((M, 256), 52).bfloat16()
    gamma, beta = 1 + 0.1 * _rand((256,), 53), 0.1 * _rand((256,), 54)
    w16 = _rand((256, 256), 55, 1 / 16).bfloat16()
    _, stats = ops.mlp_hidden_train(x, w16, _rand((256,), 56), gamma, beta)
    dout = _rand((M,), 57)
    w_out = _rand((1, 256), 58, 1 / 16).bfloat16()
    dy = torch.matmul(dout.to(torch.bfloat16)[:, None], w_out)
    want = ops.mlp_hidden_bwd(v, dy.contiguous(), stats, gamma, beta)
    got = ops.mlp_hidden_bwd_rank1(v, dout, w_out[0], stats, gamma, beta)
    for name, a, b in zip(("dv", "dgamma", "dbeta", "dbias"), got, want):
        assert torch.equal(a, b), name


def test_single_output_tower_fuses_its_last_two_layers():
    """A tower with one output (location / IoU head) runs its last hidden layer and the output Linear as one autograd node
    (``_LastHiddenOutFn``): same loss and parameter gradients as the torchvision module, like the 4-output tower above."""
    from sihl_b200 import mlp_tower
    torch.manual_seed(1)
    mlp = tvops.MLP(256, [256] * 4 + [1], norm_layer=nn.LayerNorm, activation_layer=nn.SiLU).to(DEV)
    x = _rand((2, 3000, 256), 61).requires_grad_(True)
    tgt = _rand((2, 3000, 1), 62)
    out = mlp_tower.run_tower_train(mlp, x)
    assert type(out.grad_fn.next_functions[0][0]).__name__.startswith("_LastHiddenOutFn") or "_LastHiddenOutFn" in str(out.grad_fn.next_functions)
    loss = F.mse_loss(out, tgt)
    loss.backward()
    ours = {n: p.grad.clone() for n, p in mlp.named_parameters()}
    dx = x.grad.clone()
    mlp.zero_grad(); x.grad = None
    ref_loss = F.mse_loss(mlp(x), tgt)
    ref_loss.backward()
    assert loss.item() == pytest.approx(ref_loss.item(), rel=2e-2)
    assert _cos(dx, x.grad) > 0.995 and _rel_l2(dx, x.grad) < 0.1
    for n, p in mlp.named_parameters():
        assert ours[n].shape == p.grad.shape, n
        assert _cos(ours[n], p.grad) > 0.995, (n, _cos(ours[n], p.grad))
        assert _rel_l2(ours[n], p.grad) < 0.1, (n, _rel_l2(ours[n], p.grad))


@pytest.mark.parametrize("M", [1, 5, 4099, 128 * 900])
def test_rows_colsum(M):
    rows = _rand((M, 256), 71).bfloat16()
    got = ops.rows_colsum(rows)
    want = rows.double().sum(0)
    torch.testing.assert_close(got.double(), want, rtol=1e-5, atol=1e-3 * max(1.0, M ** 0.5) * 1e-2)


def test_bn_backward_reads_its_level_in_place():
    """``ops.bn_bwd_rows`` on one level's slice of the [B, A, 256] gradient read in place == the same call on a contiguous
    copy of the slice (bit-equal), and it does not touch its neighbours' rows."""
    B, A, off, hw = 3, 700, 123, 400
    dflat = _rand((B, A, 256), 81).bfloat16()
    n = _rand((B * hw, 256), 82).bfloat16()
    scale = 1 + 0.1 * _rand((256,), 83)
    want = ops.bn_bwd_rows(dflat[:, off:off + hw].reshape(B * hw, 256).contiguous(), n, scale)
    got = ops.bn_bwd_rows(dflat, n, scale, dz_rows_per_image=A, dz_row_offset=off, rows_per_image=hw)
    for name, a, b in zip(("dy", "dgamma", "dbeta"), got, want):
        assert torch.equal(a, b), name
    with pytest.raises(ValueError):
        ops.bn_bwd_rows(dflat, n, scale, dz_rows_per_image=A, dz_row_offset=A - hw + 1, rows_per_image=hw)


def test_head_training_step_with_tensor_core_towers():
    """``mlp_backend = "tcgen05+train"``: the drop-in training step with the towers in bf16 mixed precision — loss within
    2 % of the fp32 towers, gradients of towers, laterals and inputs in the same direction."""
    model, inputs = _head(), _pyramid()
    model.train()
    for t in inputs:
        t.requires_grad_(True)
    tgt = {"classes": [torch.tensor([1, 2], device=DEV), torch.tensor([3], device=DEV)],
           "boxes": [torch.tensor([[10., 20., 100., 120.], [50., 60., 200., 220.]], device=DEV), torch.tensor([[30., 30., 90., 90.]], device=DEV)]}
    results = {}
    for backend in ("torch", "tcgen05+train"):
        model.mlp_backend = backend
        model.zero_grad()
        for t in inputs:
            t.grad = None
        loss, metrics = model.training_step(inputs, **tgt)
        loss.backward()
        results[backend] = (loss.item(), {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None},
                            inputs[3].grad.clone())
    (la, ga, ia), (lb, gb, ib) = results["torch"], results["tcgen05+train"]
    assert lb == pytest.approx(la, rel=2e-2)
    assert set(ga) == set(gb)
    for n in ga:
        if ga[n].norm() > 1e-6:
            assert _cos(ga[n], gb[n]) > 0.98, (n, _cos(ga[n], gb[n]))
    assert _cos(ia, ib) > 0.98


def test_training_laterals_on_the_tensor_core_path_match_conv_plus_batchnorm():
    """Training-mode laterals (1x1 conv + batch-statistics BatchNorm, ref :52-55, :102-105) through _LateralsTrainFn against
    the torch modules: output, every gradient (input, conv weight, BatchNorm weight and bias) and the running statistics."""
    import copy
    from sihl_b200.heads import ObjectDetection
    torch.manual_seed(2)
    model = ObjectDetection(in_channels=[3, 16, 32, 256, 256, 256], num_classes=5, num_channels=256, num_layers=1).to(DEV).train()
    with torch.no_grad():
        for lat in model.laterals:
            lat[1].weight.normal_(1, 0.2); lat[1].bias.normal_(0, 0.2)
    ref = copy.deepcopy(model)
    g = torch.Generator().manual_seed(6)
    base = [torch.randn((3, c, max(1, 96 // 2 ** l), max(1, 96 // 2 ** l)), generator=g).to(DEV) * 1.5 + 0.3 for l, c in enumerate(model.in_channels)]
    xa = [t.clone().requires_grad_(True) for t in base]
    xb = [t.clone().requires_grad_(True) for t in base]
    model.mlp_backend = "tcgen05+train"
    flat = model._flat_feats_training(xa)
    want = ref._flat_feats(xb)
    assert flat.dtype == torch.bfloat16 and flat.shape == want.shape
    torch.testing.assert_close(flat.float(), want, rtol=2e-2, atol=3e-2)
    gout = _rand(tuple(want.shape), 51)
    flat.backward(gout.bfloat16())
    want.backward(gout)
    for lvl in (3, 4, 5):
        assert _cos(xa[lvl].grad, xb[lvl].grad) > 0.995 and _rel_l2(xa[lvl].grad, xb[lvl].grad) < 0.1, lvl
    for (na, pa), (nb, pb) in zip(model.laterals.named_parameters(), ref.laterals.named_parameters()):
        assert na == nb and _cos(pa.grad, pb.grad) > 0.995 and _rel_l2(pa.grad, pb.grad) < 0.1, (na, _cos(pa.grad, pb.grad))
    for la, lb in zip(model.laterals, ref.laterals):
        torch.testing.assert_close(la[1].running_mean, lb[1].running_mean, rtol=2e-2, atol=2e-3)
        torch.testing.assert_close(la[1].running_var, lb[1].running_var, rtol=2e-2, atol=2e-3)
        assert int(la[1].num_batches_tracked) == int(lb[1].num_batches_tracked) == 1


def test_head_training_step_with_tensor_core_laterals_and_towers():
    """The whole drop-in training step with laterals AND towers on the tensor-core path (256-channel pyramid) against the
    torch modules: loss within 2 %, parameter gradients in the same direction."""
    import copy
    from sihl_b200.heads import ObjectDetection
    torch.manual_seed(4)
    model = ObjectDetection(in_channels=[3, 16, 32, 256, 256, 256], num_classes=6, num_channels=256, num_layers=2, max_instances=20).to(DEV).train()
    ref = copy.deepcopy(model)
    model.mlp_backend = "tcgen05+train"
    g = torch.Generator().manual_seed(8)
    inputs = [torch.randn((2, c, max(1, 128 // 2 ** l), max(1, 128 // 2 ** l)), generator=g).to(DEV) for l, c in enumerate(model.in_channels)]
    tgt = {"classes": [torch.tensor([1, 2], device=DEV), torch.tensor([3], device=DEV)],
           "boxes": [torch.tensor([[10., 20., 100., 120.], [50., 60., 110., 90.]], device=DEV), torch.tensor([[30., 30., 90., 90.]], device=DEV)]}
    loss, _ = model.training_step(inputs, **tgt)
    loss.backward()
    ref_loss, _ = ref.training_step(inputs, **tgt)
    ref_loss.backward()
    assert loss.item() == pytest.approx(ref_loss.item(), rel=2e-2)
    checked = 0
    for (n, p), (_, q) in zip(model.named_parameters(), ref.named_parameters()):
        if q.grad is not None and q.grad.norm() > 1e-6:
            assert p.grad is not None and _cos(p.grad, q.grad) > 0.97, (n, _cos(p.grad, q.grad))
            checked += 1
    assert checked > 20


def test_optimizer_trajectory_with_the_tensor_core_training_path():
    """Fifteen AdamW steps on one batch, torch modules vs ``mlp_backend = "tcgen05+train"`` from the same initial weights: both
    losses fall and stay within a few percent of each other (bf16 mixed precision vs fp32), running statistics follow."""
    import copy
    from sihl_b200.heads import ObjectDetection
    torch.manual_seed(7)
    a = ObjectDetection(in_channels=[3, 16, 32, 256, 256, 256], num_classes=6, num_channels=256, num_layers=2, max_instances=20).to(DEV).train()
    b = copy.deepcopy(a)
    b.mlp_backend = "tcgen05+train"
    g = torch.Generator().manual_seed(9)
    inputs = [torch.randn((2, c, max(1, 128 // 2 ** l), max(1, 128 // 2 ** l)), generator=g).to(DEV) for l, c in enumerate(a.in_channels)]
    tgt = {"classes": [torch.tensor([1, 2], device=DEV), torch.tensor([3], device=DEV)],
           "boxes": [torch.tensor([[10., 20., 100., 120.], [50., 60., 110., 90.]], device=DEV), torch.tensor([[30., 30., 90., 90.]], device=DEV)]}
    curves = []
    for model in (a, b):
        opt = torch.optim.AdamW(model.parameters(), lr=2e-4)
        losses = []
        for _ in range(15):
            opt.zero_grad(set_to_none=True)
            loss, _ = model.training_step(inputs, **tgt)
            loss.backward()
            opt.step()
            losses.append(loss.item())
        curves.append(losses)
    la, lb = curves
    assert la[-1] < 0.9 * la[0] and lb[-1] < 0.9 * lb[0], (la, lb)
    for x, y in zip(la, lb):
        assert y == pytest.approx(x, rel=0.06), (la, lb)
    for ma, mb in zip(a.laterals, b.laterals):
        torch.testing.assert_close(mb[1].running_mean, ma[1].running_mean, rtol=5e-2, atol=5e-3)
        torch.testing.assert_close(mb[1].running_var, ma[1].running_var, rtol=5e-2, atol=5e-3)

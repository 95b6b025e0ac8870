"""Row N4 host logic without a GPU: tower structure recognition, padding rules, C-ABI argument checks."""
import ctypes

import pytest
import torch
from torch import nn
from torchvision import ops as tvops

from sihl_b200 import _native, ops
from sihl_b200.mlp_tower import PackedTower, _split


def _tower(width=256, layers=4, out=80):
    return tvops.MLP(width, [width] * layers + [out], norm_layer=nn.LayerNorm, activation_layer=nn.SiLU)


def test_recognises_the_reference_tower_and_nothing_else():
    """ref src/sihl/heads/object_detection.py:51, :56-60: Linear -> LayerNorm -> SiLU -> Dropout, x num_layers, + Linear."""
    hidden, last = _split(_tower())
    assert len(hidden) == 4 and last.out_features == 80
    assert PackedTower.supported(_tower()) and PackedTower.supported(_tower(layers=0, out=1))
    assert not PackedTower.supported(_tower(width=128))
    assert not PackedTower.supported(tvops.MLP(256, [256, 4], norm_layer=nn.LayerNorm, activation_layer=nn.ReLU))
    assert not PackedTower.supported(tvops.MLP(256, [256, 4], norm_layer=None, activation_layer=nn.SiLU))
    assert not PackedTower.supported(tvops.MLP(256, [256, 4], norm_layer=nn.LayerNorm, activation_layer=nn.SiLU, bias=False))
    assert not PackedTower.supported(_tower(out=300))
    assert not PackedTower.supported(nn.Sequential(nn.Linear(256, 256), nn.SiLU()))


def test_packing_pads_the_output_layer_and_follows_parameter_updates():
    mlp = _tower(out=5)
    packed = PackedTower(mlp).refresh()
    w, b, n_out = packed.out
    assert n_out == 5 and tuple(w.shape) == (16, 256) and w.dtype == torch.bfloat16 and (w[5:] == 0).all() and (b[5:] == 0).all()
    assert len(packed.hidden) == 4 and packed.hidden[0][0].dtype == torch.bfloat16 and packed.hidden[0][4] == pytest.approx(1e-5)
    first = packed.hidden[0][0]
    assert packed.refresh().hidden[0][0] is first                      # nothing changed: no re-pack
    with torch.no_grad():
        mlp[0].weight.mul_(2.0)
    assert packed.refresh().hidden[0][0] is not first
    torch.testing.assert_close(packed.hidden[0][0].float(), mlp[0].weight.detach().bfloat16().float())


@pytest.mark.parametrize("out_features,pad", [(1, 16), (4, 16), (16, 16), (17, 32), (80, 96), (97, 128), (129, 256), (256, 256)])
def test_output_padding_rule(out_features, pad):
    assert ops.mlp_out_pad(out_features) == pad


def test_cabi_rejects_bad_arguments_before_touching_the_device():
    lib = _native.load()
    buf = (ctypes.c_char * 64)()
    p = ctypes.addressof(buf)
    assert lib.sihl_od_mlp_hidden(p, 128, 128, p, p, p, p, 1e-5, p, None) == 1          # channels != 256
    assert lib.sihl_od_mlp_hidden(p, -1, 256, p, p, p, p, 1e-5, p, None) == 1
    assert lib.sihl_od_mlp_hidden(None, 128, 256, p, p, p, p, 1e-5, p, None) == 1
    assert lib.sihl_od_mlp_hidden(p + 8, 128, 256, p, p, p, p, 1e-5, p, None) == 1      # 16-byte alignment
    assert lib.sihl_od_mlp_hidden(None, 0, 256, None, None, None, None, 1e-5, None, None) == 0   # empty input: nothing to do
    assert lib.sihl_od_mlp_out(p, 128, 256, p, p, 24, 4, p, None) == 1                  # n_pad not a supported width
    assert lib.sihl_od_mlp_out(p, 128, 256, p, p, 16, 17, p, None) == 1                 # out_cols > n_pad
    assert lib.sihl_od_mlp_out(p, 128, 64, p, p, 16, 4, p, None) == 1
    assert lib.sihl_od_mlp_out(None, 0, 256, None, None, 16, 4, None, None) == 0


def test_head_backend_switch_is_validated():
    from sihl_b200.heads import ObjectDetection
    model = ObjectDetection(in_channels=[3] + [8] * 5, num_classes=3, num_channels=16, num_layers=1)
    assert model.mlp_backend == "torch"
    x = torch.zeros((1, 4, 16))
    assert model._use_tcgen05(x) is False
    model.mlp_backend = "tcgen05"
    assert model._use_tcgen05(x) is False                              # CPU tensor: the torch modules
    model.mlp_backend = "cutlass"
    with pytest.raises(ValueError):
        model._use_tcgen05(x)


def test_cabi_training_and_lateral_entries_reject_bad_arguments():
    lib = _native.load()
    buf = (ctypes.c_char * 64)()
    p = ctypes.addressof(buf)
    assert lib.sihl_od_mlp_hidden_train(p, 128, 256, p, p, p, p, 1e-5, p, None, None, None) == 1  # row_stats missing
    assert lib.sihl_od_mlp_hidden_train(p, 128, 64, p, p, p, p, 1e-5, p, p, None, None) == 1
    assert lib.sihl_od_mlp_hidden_train(p, 128, 256, p, p, p, p, 1e-5, p, p, p + 2, None) == 1     # v_bf16 misaligned
    assert lib.sihl_od_mlp_hidden_train(None, 0, 256, None, None, None, None, 1e-5, None, None, None, None) == 0
    assert lib.sihl_od_mlp_hidden_bwd(p, p, p, p, p, 128, 256, p, p, 0, None) == 1                 # no partial rows
    assert lib.sihl_od_mlp_hidden_bwd_rank1(p, p, None, p, p, p, 128, 256, p, p, 8, None) == 1     # w_out missing
    assert lib.sihl_od_bn_bwd_colsums_map(p, 64, 100, 50, p, 128, 256, p, 8, None) == 1            # slice leaves the image (or no GPU)
    assert lib.sihl_od_bn_bwd_apply_map(p, 64, 100, 50, p, p, p, p, 128, 256, p, None) == 1
    assert lib.sihl_od_bn_bwd_apply_map(None, 1, 1, 0, None, None, None, None, 0, 256, None, None) == 0
    assert lib.sihl_od_rows_colsum(p, 128, 128, p, 8, None) == 1                                   # channels
    assert lib.sihl_od_mlp_hidden_bwd(p, p, p, p, p, 128, 128, p, p, 8, None) == 1
    assert lib.sihl_od_bf16_to_f32(p, 12, p, None) == 1                                            # n % 8
    assert lib.sihl_od_bf16_to_f32(p + 2, 16, p, None) == 1                                        # alignment
    assert lib.sihl_od_bf16_to_f32(None, 0, None, None) == 0
    assert lib.sihl_od_lateral_rows(p, 2, 100, 64, p, None) == 1                                   # channels % 64
    assert lib.sihl_od_lateral_rows(p, 0, 256, 64, p, None) == 0
    assert lib.sihl_od_lateral_linear(p, 128, 256, p, p, 0, 128, 0, p, None) == 1                  # rows_per_image <= 0
    assert lib.sihl_od_lateral_linear(p, 128, 256, p, p, 64, 100, 50, p, None) == 1                # slice leaves the image
    assert lib.sihl_od_lateral_linear(p, 128, 128, p, p, 64, 64, 0, p, None) == 1


def test_training_backend_is_only_used_when_gradients_are_recorded():
    from sihl_b200.heads import ObjectDetection
    model = ObjectDetection(in_channels=[3] + [8] * 5, num_classes=3, num_channels=16, num_layers=1)
    model.mlp_backend = "tcgen05+train"
    x = torch.zeros((1, 4, 16))
    assert model._use_tcgen05_training(x) is False                     # CPU tensor
    assert model._use_tcgen05(x) is False
    out = model._tower("loc_head", x)                                  # -> the torch module
    assert out.shape == (1, 4, 1) and out.requires_grad


def test_fan_out_node_matches_plain_autograd():
    """``_FanOutFn`` (one autograd node for the three consumers of the tower features) against what autograd does by itself:
    same outputs, same input gradient — including repeated index entries (the padding rows repeat row 0) and consumers
    that receive no gradient."""
    from sihl_b200.mlp_tower import _FanOutFn
    g = torch.Generator().manual_seed(0)
    x0 = torch.randn((50, 8), generator=g, dtype=torch.float64)
    idx = torch.tensor([3, 7, 49, 0, 0, 0], dtype=torch.int32)
    wa, wb, wr = (torch.randn(s, generator=g, dtype=torch.float64) for s in ((50, 8), (50, 8), (6, 8)))
    for use in ((True, True, True), (True, False, True), (False, False, True), (True, True, False)):
        x = x0.clone().requires_grad_(True)
        a, b, rows = _FanOutFn.apply(x, idx)
        assert torch.equal(a, x0) and torch.equal(b, x0) and torch.equal(rows, x0[idx.long()])
        loss = sum(((t * w).sum() for t, w, u in zip((a, b, rows), (wa, wb, wr), use) if u))
        loss.backward()
        y = x0.clone().requires_grad_(True)
        ref = sum(((t * w).sum() for t, w, u in zip((y, y, y.index_select(0, idx)), (wa, wb, wr), use) if u))
        ref.backward()
        torch.testing.assert_close(x.grad, y.grad, rtol=0, atol=1e-12)

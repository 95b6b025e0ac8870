"""ncu target for row N4's training kernels: a few launches of the LayerNorm + SiLU backward (both variants) and of the
training forward at M = 537 600."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sihl_b200 import ops
dev = torch.device("cuda", 0)
M = 64 * 8400
x = torch.randn((M, 256), device=dev).bfloat16()
dy = torch.randn((M, 256), device=dev).bfloat16()
w = (torch.randn((256, 256), device=dev) / 16).bfloat16()
b = torch.zeros(256, device=dev); g = torch.ones(256, device=dev)
dout = torch.randn((M,), device=dev); wo = (torch.randn((256,), device=dev) / 16).bfloat16()
for _ in range(3):
    y, st, v = ops.mlp_hidden_train(x, w, b, g, b, save_pre=True)
    ops.mlp_hidden_bwd(v, dy, st, g, b)
    ops.mlp_hidden_bwd_rank1(v, dout, wo, st, g, b)
torch.cuda.synchronize()

"""ncu target for row N4: a few launches of one hidden layer and one 80-class output layer at M = 545 600."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sihl_b200 import ops
dev = torch.device("cuda", 0)
M = 64 * 8525
x = torch.randn((M, 256), device=dev).bfloat16(); y = torch.empty_like(x)
w = (torch.randn((256, 256), device=dev) / 16).bfloat16()
b = torch.zeros(256, device=dev); g = torch.ones(256, device=dev)
wo = torch.zeros((96, 256), dtype=torch.bfloat16, device=dev); bo = torch.zeros(96, device=dev)
for _ in range(4):
    ops.mlp_hidden(x, w, b, g, b, out=y)
    ops.mlp_out(x, wo, bo, 80)
torch.cuda.synchronize()

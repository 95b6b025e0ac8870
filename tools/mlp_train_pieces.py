"""Developer probe: time the pieces of one hidden layer's training backward at M = 545 600."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sihl_b200 import ops
dev = torch.device("cuda", 0)
M = 64 * 8525
def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
x = torch.randn((M, 256), device=dev).bfloat16(); dy = torch.randn((M, 256), device=dev).bfloat16()
w = (torch.randn((256, 256), device=dev) / 16).bfloat16(); b = torch.zeros(256, device=dev); g = torch.ones(256, device=dev)
y, stats = ops.mlp_hidden_train(x, w, b, g, b)
v = ops.linear_bf16(x, w, b)
dv, *_ = ops.mlp_hidden_bwd(v, dy, stats, g, b)
print("fwd train        ", timeit(lambda: ops.mlp_hidden_train(x, w, b, g, b)))
print("recompute linear ", timeit(lambda: ops.linear_bf16(x, w, b)))
print("bwd rows         ", timeit(lambda: ops.mlp_hidden_bwd(v, dy, stats, g, b)))
print("dx linear        ", timeit(lambda: ops.linear_bf16(dv, w.t().contiguous(), torch.zeros_like(b))))
print("dW matmul dv.t@x ", timeit(lambda: torch.matmul(dv.t(), x)))
print("dW via x.t@dv .t ", timeit(lambda: torch.matmul(x.t(), dv)))
xf = x.float()
print("x fp32->bf16     ", timeit(lambda: xf.to(torch.bfloat16)))
print("dx bf16->fp32    ", timeit(lambda: dv.float()))
d1 = torch.randn((M, 1), device=dev)
w1 = w[:1]
print("dx bf16->fp32 own", timeit(lambda: ops.bf16_to_f32(dv)))
print("out bwd dx       ", timeit(lambda: torch.matmul(d1.to(torch.bfloat16), w1)))
print("out bwd dW       ", timeit(lambda: torch.matmul(d1.to(torch.bfloat16).t(), x)))

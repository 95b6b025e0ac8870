"""Developer probe for row N4 (sihl_od_mlp_*): one case per process (a device trap poisons the CUDA context).
python tools/mlp_debug.py out16|out256|hidden [M]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sihl_b200 import ops

case = sys.argv[1]
M = int(sys.argv[2]) if len(sys.argv) > 2 else 128
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(5)
x = torch.randn((M, 256), generator=g, device=dev).bfloat16()
if case.startswith("out"):
    n = int(case[3:])
    w = (torch.randn((n, 256), generator=g, device=dev) / 16).bfloat16()
    b = torch.randn((n,), generator=g, device=dev)
    y = ops.mlp_out(x, w, b, n)
    torch.cuda.synchronize()
    ref = (x.double() @ w.double().T + b.double()).float()
    err = (y - ref).abs()
    print(case, "M", M, "max abs err", float(err.max()), "ref max", float(ref.abs().max()))
    if err.max() > 1e-2:
        bad = (err > 1e-2)
        print("bad fraction", float(bad.float().mean()), "bad rows", bad.any(1).nonzero().flatten()[:16].tolist(), "bad cols", bad.any(0).nonzero().flatten()[:32].tolist())
        print("ours[0,:8]", y[0, :8].tolist()); print("ref [0,:8]", ref[0, :8].tolist())
        # which K chunks contribute?  partial products per 64-wide K chunk and per 16-wide K step
        for c in range(4):
            part = (x[:, c*64:(c+1)*64].double() @ w[:, c*64:(c+1)*64].double().T).float()
            print("chunk", c, "partial[0,:4]", part[0, :4].tolist())
else:
    w = (torch.randn((256, 256), generator=g, device=dev) / 16).bfloat16()
    b = torch.randn((256,), generator=g, device=dev)
    gamma = 1 + 0.1 * torch.randn((256,), generator=g, device=dev)
    beta = 0.1 * torch.randn((256,), generator=g, device=dev)
    y = ops.mlp_hidden(x, w, b, gamma, beta)
    torch.cuda.synchronize()
    pre = (x.double() @ w.double().T + b.double()).float()
    ref = torch.nn.functional.silu(torch.nn.functional.layer_norm(pre, (256,), gamma, beta, 1e-5))
    err = (y.float() - ref).abs()
    print(case, "M", M, "max abs err", float(err.max()), "mean abs err", float(err.mean()), "ref max", float(ref.abs().max()))
    print("ours[0,:8]", y[0, :8].float().tolist()); print("ref [0,:8]", ref[0, :8].tolist())

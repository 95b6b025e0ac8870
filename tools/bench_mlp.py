"""Row N4 measurement (developer tool; `python tools/bench_mlp.py [--m 545600] > profiles/r02_n4_mlp.json`):
one hidden layer (Linear 256->256 + LayerNorm + SiLU), one output layer and the four whole towers of the head at the
cfg1 location count (640^2, batch 64: M = 545 600), through the tcgen05 kernels and through torch eager (fp32 = what
the reference runs, and bf16 autocast).  Inputs rotate over 3 buffers (3 x 279 MB > L2); CUDA events; warm-up first."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch import nn
from torchvision import ops as tvops
from sihl_b200 import ops
from sihl_b200.mlp_tower import PackedTower, run_tower

ap = argparse.ArgumentParser()
ap.add_argument("--m", type=int, default=64 * 8525)
ap.add_argument("--iters", type=int, default=30)
ap.add_argument("--classes", type=int, default=80)
args = ap.parse_args()
dev = torch.device("cuda", 0)
M = args.m
peaks = {}
try:
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
except Exception:
    pass
HBM = float(peaks.get("hbm_gbs", 6471.1)); TF = float(peaks.get("bf16_tflops", 1660.7))

def timeit(fn, iters=args.iters, warm=5):
    for i in range(warm): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

torch.manual_seed(0)
xs = [torch.randn((M, 256), device=dev).bfloat16() for _ in range(3)]
ys = [torch.empty((M, 256), dtype=torch.bfloat16, device=dev) for _ in range(3)]
lin, ln = nn.Linear(256, 256).to(dev), nn.LayerNorm(256).to(dev)
w, b, g, be = lin.weight.detach().bfloat16().contiguous(), lin.bias.detach().float(), ln.weight.detach().float(), ln.bias.detach().float()
res = {"M": M, "peaks": {"hbm_gbs": HBM, "bf16_tflops": TF, "source": "MEASURED_PEAKS.json" if peaks else "fallback"}}

ms = timeit(lambda i: ops.mlp_hidden(xs[i % 3], w, b, g, be, out=ys[i % 3]))
flops, byts = 2.0 * M * 256 * 256, M * 1024.0
res["hidden_layer"] = {"ms": ms, "tflops": flops / ms / 1e9, "gbs": byts / ms / 1e6, "frac_of_hbm_peak": byts / ms / 1e6 / HBM,
                       "frac_of_tensor_peak": flops / ms / 1e9 / TF, "algorithmic_bytes": byts, "flops": flops}
with torch.no_grad():
    xf = [x.float() for x in xs[:2]]
    seq = nn.Sequential(lin, ln, nn.SiLU())
    res["hidden_layer"]["torch_eager_fp32_ms"] = timeit(lambda i: seq(xf[i % 2]), iters=10, warm=3)
    def bf(i):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            return seq(xs[i % 3])
    res["hidden_layer"]["torch_eager_bf16_autocast_ms"] = timeit(bf, iters=10, warm=3)
    seq_b = nn.Sequential(lin, ln, nn.SiLU()).bfloat16() if False else None
    del xf
torch.cuda.empty_cache()

for n_out in (1, 4, args.classes):
    n_pad = ops.mlp_out_pad(n_out)
    wo = torch.zeros((n_pad, 256), dtype=torch.bfloat16, device=dev); wo[:n_out] = torch.randn((n_out, 256), device=dev).bfloat16() / 16
    bo = torch.zeros((n_pad,), device=dev)
    outs = [torch.empty((M, n_out), device=dev) for _ in range(3)]
    ms = timeit(lambda i: ops.mlp_out(xs[i % 3], wo, bo, n_out, out=outs[i % 3]))
    byts = M * (512.0 + 4 * n_out)
    res[f"output_layer_{n_out}"] = {"ms": ms, "gbs": byts / ms / 1e6, "frac_of_hbm_peak": byts / ms / 1e6 / HBM, "n_pad": n_pad}
    del outs

# the four towers of the head (loc, iou: 1; box: 4; cls: classes), 4 hidden layers each
towers = {n: tvops.MLP(256, [256] * 4 + [o], norm_layer=nn.LayerNorm, activation_layer=nn.SiLU).to(dev).eval()
          for n, o in (("loc", 1), ("iou", 1), ("box", 4), ("cls", args.classes))}
packed = {n: PackedTower(m).refresh() for n, m in towers.items()}
scratch = (ys[0], ys[1])
def ours(i):
    for n in towers: run_tower(packed[n], xs[i % 3], scratch)
ms = timeit(ours, iters=10, warm=3)
flops = sum(2.0 * M * 256 * (256 * 4 + ops.mlp_out_pad(o)) for o in (1, 1, 4, args.classes))
res["four_towers"] = {"ms": ms, "tflops": flops / ms / 1e9, "frac_of_tensor_peak": flops / ms / 1e9 / TF, "launches": 4 * 5}
del ys
torch.cuda.empty_cache()
with torch.no_grad():
    try:
        xf = xs[0].float()
        def ref(i):
            for n in towers: towers[n](xf)
        res["four_towers"]["torch_eager_fp32_ms"] = timeit(ref, iters=3, warm=1)
        del xf
    except torch.OutOfMemoryError as e:
        res["four_towers"]["torch_eager_fp32_ms"] = None
    torch.cuda.empty_cache()
    def refb(i):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            for n in towers: towers[n](xs[i % 3])
    res["four_towers"]["torch_eager_bf16_autocast_ms"] = timeit(refb, iters=3, warm=1)

# training: one tower (loc: 4 hidden + 1 output column) forward + backward over M locations
from sihl_b200.mlp_tower import run_tower_train
tower = towers["loc"].train()
xt = xs[0].float().requires_grad_(True)
gy = torch.randn((M, 1), device=dev)
def train_ours(i):
    tower.zero_grad(set_to_none=True); xt.grad = None
    run_tower_train(tower, xt).backward(gy)
def train_torch(i):
    tower.zero_grad(set_to_none=True); xt.grad = None
    tower(xt).backward(gy)
def train_autocast(i):
    tower.zero_grad(set_to_none=True); xt.grad = None
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = tower(xt)
    out.backward(gy.to(out.dtype))
res["tower_forward_backward"] = {"ms": timeit(train_ours, iters=5, warm=2)}
torch.cuda.empty_cache()
try:
    res["tower_forward_backward"]["torch_eager_fp32_ms"] = timeit(train_torch, iters=3, warm=1)
except torch.OutOfMemoryError:
    res["tower_forward_backward"]["torch_eager_fp32_ms"] = None
torch.cuda.empty_cache()
res["tower_forward_backward"]["torch_eager_bf16_autocast_ms"] = timeit(train_autocast, iters=3, warm=1)
del xt, gy
tower.eval()
torch.cuda.empty_cache()

# the head's forward (ref :99-122) at cfg1: laterals + loc tower on every location + top-K + cls / box towers on K rows + decode
del xs, towers, packed
torch.cuda.empty_cache()
from sihl_b200.heads import ObjectDetection
B, size = 64, 640
model = ObjectDetection(in_channels=[3, 64, 128, 256, 256, 256], num_classes=args.classes, num_channels=256, num_layers=4).to(dev).eval()
g = torch.Generator(device=dev); g.manual_seed(0)
inputs = [torch.randn((B, c, max(1, size // 2 ** l), max(1, size // 2 ** l)), generator=g, device=dev) if l >= 3 or l == 0 else torch.empty((B, c, 1, 1), device=dev)
          for l, c in enumerate(model.in_channels)]
with torch.no_grad():
    model.mlp_backend = "torch"
    t_torch = timeit(lambda i: model.forward(inputs), iters=5, warm=2)
    model.mlp_backend = "tcgen05"
    t_tc = timeit(lambda i: model.forward(inputs), iters=10, warm=3)
    feats = model._flat_feats(inputs)
    t_lat = timeit(lambda i: model._flat_feats(inputs), iters=10, warm=3)
    t_lat_tc = timeit(lambda i: model._tower_feats(inputs), iters=10, warm=3)
    lvl = inputs[3]
    rows = ops.lateral_rows(lvl)
    t_rows = timeit(lambda i: ops.lateral_rows(lvl, out=rows), iters=20, warm=3)
    wt, bias = model._folded_laterals()[0]
    flat_tc = model._tower_feats(inputs)
    t_lin = timeit(lambda i: ops.lateral_linear(rows, wt, bias, lvl.shape[2] * lvl.shape[3], flat_tc, 0), iters=20, warm=3)
res["head_forward_cfg1"] = {"batch": B, "image": size, "locations": int(feats.shape[1]), "torch_towers_ms": t_torch, "tcgen05_towers_ms": t_tc,
                            "laterals_and_concat_ms": t_lat, "speedup": t_torch / t_tc,
                            "laterals_tensor_cores_ms": t_lat_tc,
                            "level3_rows_kernel": {"ms": t_rows, "gbs": lvl.numel() * 6 / t_rows / 1e6, "frac_of_hbm_peak": lvl.numel() * 6 / t_rows / 1e6 / HBM},
                            "level3_linear_kernel": {"ms": t_lin, "gbs": rows.numel() * 4 / t_lin / 1e6, "frac_of_hbm_peak": rows.numel() * 4 / t_lin / 1e6 / HBM}}

# the drop-in head's whole training step (ref :124-217) + backward at cfg1: laterals (torch) + four towers + the hot path
from sihl_b200 import synth
gt = synth.gt_batch_np(3, B, size, size, args.classes, 100)
tb = [torch.from_numpy(b_).to(dev) for b_, _ in gt.per_image()]
tc = [torch.from_numpy(c_).to(dev) for _, c_ in gt.per_image()]
model.train()
def train_step(i):
    model.zero_grad(set_to_none=True)
    loss, _ = model.training_step(inputs, classes=tc, boxes=tb)
    loss.backward()
res["head_training_step_cfg1"] = {}
for backend in ("tcgen05+train", "torch"):
    model.mlp_backend = backend
    torch.cuda.empty_cache()
    res["head_training_step_cfg1"][backend.replace("+", "_") + "_ms"] = timeit(train_step, iters=3 if backend == "torch" else 6, warm=2)
model.zero_grad(set_to_none=True)
model.eval()
torch.cuda.empty_cache()

# small-batch serving: eager (one host launch per kernel) vs one CUDA graph per call
from sihl_b200.serving import GraphedInference
del inputs, feats, rows, flat_tc
torch.cuda.empty_cache()
res["serving_latency_ms"] = {}
for b in (1, 8):
    small = [torch.randn((b, c, max(1, size // 2 ** l), max(1, size // 2 ** l)), generator=g, device=dev) if l >= 3 or l == 0 else torch.empty((b, c, 1, 1), device=dev)
             for l, c in enumerate(model.in_channels)]
    entry = {}
    with torch.no_grad():
        for backend in ("torch", "tcgen05"):
            model.mlp_backend = backend
            entry[f"{backend}_towers_eager"] = timeit(lambda i: model.forward(small), iters=30, warm=5)
            graphed = GraphedInference(model.forward, small)
            entry[f"{backend}_towers_graph"] = timeit(lambda i: graphed(small), iters=100, warm=10)
            del graphed
    res["serving_latency_ms"][f"batch_{b}"] = entry
print(json.dumps(res, indent=1))

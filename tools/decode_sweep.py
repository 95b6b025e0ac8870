"""Developer probe: k_dense_decode_tma alone under different ring shapes (rows per stage x stages x CTAs per SM), for a
workload of bench.py.  Each launch gets fresh zeroed candidate counters; inputs rotate over 3 sets (> L2 at cfg1).
    python tools/decode_sweep.py crowd > gpurun_out/decode_sweep_crowd.json"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sihl_b200 import ops, synth
from bench import WORKLOADS, measured_peaks

name = sys.argv[1] if len(sys.argv) > 1 else "crowd"
w = WORKLOADS[name]
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
H, W, B, C = w["height"], w["width"], w["batch"], w["classes"]
levels = synth.level_sizes(H, W)
A = synth.num_anchors(levels)
gen = torch.Generator(device=dev); gen.manual_seed(1234)
n_sets = min(16, max(3, -(-(400 << 20) // (B * A * (C + 5) * 4))))          # rotating inputs: together > 3x the 126 MB L2
sets = [synth.dense_maps_torch(gen, B, A, C, dev) for _ in range(n_sets)]
off, sc, _ = ops.anchor_tables(levels, W, H, dev)
cand = ops.CandidateBuffers.allocate(B, A, dev)
iters = 400
scratch = torch.zeros((iters + 6, B), dtype=torch.int32, device=dev)
flush = None
peak, _ = measured_peaks()


def run(i):
    loc, iou, box, cls = sets[i % n_sets]
    cand.count = scratch[i]
    ops.dense_decode(loc, cls, box, off, sc, W, H, 0.05, cand, zero_counts=False)


def timed_graph():
    """SWEEP_GRAPH=1: 100 back-to-back launches captured into one CUDA graph, replayed 4x — no host time per launch (the
    eager loop below costs ~14 us of Python + ctypes per call, which is what a 9-14 us kernel then measures)."""
    scratch.zero_()
    for i in range(6): run(iters + i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        scratch[:100].zero_()
        for i in range(100): run(i)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters // 100): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def timed():
    if os.environ.get("SWEEP_GRAPH"):
        return timed_graph()
    scratch.zero_()
    for i in range(6): run(iters + i)
    torch.cuda.synchronize()
    if flush is None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters): run(i)
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters
    total = 0.0                                       # small maps: flush L2 between launches, time each launch
    for i in range(iters):
        flush.fill_(i & 255)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(i); e1.record(); torch.cuda.synchronize()
        total += e0.elapsed_time(e1)
    return total / iters


rows = []
timed()                                                # clocks up before the first measured configuration
CONFIGS = [(None, None, None)] + [(r, s, c) for r in (64, 32) for s in (2, 3, 4, 6, 8) for c in (1, 2, 3, 4, 6)] + [(None, None, None)]
if os.environ.get("SWEEP_SHORT"):
    CONFIGS = [(None, None, None)] * 4
for r, s, c in CONFIGS:
    for k, v in (("SIHL_DECODE_ROWS", r), ("SIHL_DECODE_STAGES", s), ("SIHL_DECODE_CTAS_PER_SM", c)):
        if v is None: os.environ.pop(k, None)
        else: os.environ[k] = str(v)
    if r is not None and r * C * 4 * s * c > 200 * 1024:
        continue
    ms = timed()
    cand_mean = float(scratch[:100 if os.environ.get("SWEEP_GRAPH") else iters].float().mean().item())
    bytes_ = B * 4 * A * (C + 1) + B * cand_mean * 44
    rows.append({"rows": r, "stages": s, "ctas_per_sm": c, "us": ms * 1e3, "gbs": bytes_ / ms / 1e6, "frac": bytes_ / ms / 1e6 / peak})
    print(rows[-1], file=sys.stderr)
print(json.dumps({"workload": name, "rotating_input_sets": n_sets, "rows": rows}, indent=1))

"""Developer probe (not part of the product): time the two chains of the pipeline separately
and together with CUDA-graph replay, rotating input sets.  python tools_chain_times.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sihl_b200 import ops, synth
from sihl_b200.pipeline import DetectionHeadPipeline, StepInputs

dev = torch.device("cuda", 0)
H = W = 640; B, C, G, K = 64, 80, 100, 100
levels = synth.level_sizes(H, W)
pipe = DetectionHeadPipeline(levels, W, H, B, C, B * G, dev)
gen = torch.Generator(device=dev); gen.manual_seed(1234)
sets, outs = [], []
for _ in range(3):
    boxes, classes, offsets = synth.gt_batch_torch(gen, B, H, W, C, G, dev)
    loc, iou, box, cls = synth.dense_maps_torch(gen, B, pipe.A, C, dev)
    sets.append(StepInputs(loc, iou, box, cls, ops.GtBatch(boxes, classes, offsets, [G] * B)))
    outs.append(pipe.new_outputs())
torch.cuda.synchronize()

def capture(fn):
    gs = []
    for i in range(3):
        warm = torch.cuda.Stream(device=dev); warm.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(warm): fn(i)
        torch.cuda.current_stream().wait_stream(warm); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g): fn(i)
        gs.append(g)
    return gs

def timeit(gs, n=600):
    for s in range(60): gs[s % 3].replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(n): gs[s % 3].replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

lib = pipe.lib
def k_select(i):
    x, out = sets[i], outs[i]; gt = x.gt; p = ops._p; st = torch.cuda.current_stream().cuda_stream
    lib.sihl_od_assign_select(p(pipe.anchors), p(pipe.terms), pipe.A, pipe._hw.ctypes.data, len(pipe._hw), W, H, p(gt.boxes), p(gt.offsets), B, gt.total, 9, p(pipe.sel_anchor), p(pipe.sel_val), p(pipe.best_iou), p(out.sums), st)
def k_resolve(i):
    x, out = sets[i], outs[i]; gt = x.gt; p = ops._p; st = torch.cuda.current_stream().cuda_stream
    lib.sihl_od_assign_resolve(p(pipe.sel_anchor), p(pipe.sel_val), p(pipe.best_iou), p(gt.offsets), B, pipe.A, 9, 1, p(x.loc_logits), p(x.iou_preds), p(out.assignment), p(out.rel_iou), p(out.sums), p(pipe.tile_pos_count), p(pipe.tile_pos_rows), p(x.box_raw), p(x.cls_logits), C, p(pipe.pos_chunks), p(pipe.tile_pos_aux), st)
def k_pos(i):
    x, out = sets[i], outs[i]; gt = x.gt; p = ops._p; st = torch.cuda.current_stream().cuda_stream
    lib.sihl_od_pos_loss_tiles(p(pipe.pos_chunks), p(pipe.tile_pos_rows), p(pipe.tile_pos_aux), B, pipe.A, p(pipe.offsets), p(pipe.scales), W, H, p(gt.boxes), p(gt.classes), p(gt.offsets), p(x.box_raw), p(x.cls_logits), C, p(out.sums), p(out.losses), p(pipe.done_counter), st)
def k_decode(i):
    x = sets[i]
    ops.dense_decode(x.loc_logits, x.cls_logits, x.box_raw, pipe.offsets, pipe.scales, W, H, 0.05, pipe.cand, zero_counts=False)
def k_nms(i):
    out = outs[i]
    ops.nms_topk(pipe.cand, B, 0.5, K, (out.num_instances, out.scores, out.classes, out.boxes), reset_counts=True)

res = {}
res["train_chain"] = timeit(capture(lambda i: pipe.train_chain(sets[i], outs[i])))
res["infer_chain"] = timeit(capture(lambda i: pipe.infer_chain(sets[i], outs[i])))
res["both_2streams"] = timeit(capture(lambda i: pipe.step(sets[i], outs[i])))
lo, hi = torch.cuda.Stream.priority_range() if hasattr(torch.cuda.Stream, "priority_range") else (0, -1)
pipe.side = torch.cuda.Stream(device=dev, priority=0)
hp = torch.cuda.Stream(device=dev, priority=-1)
def step_hp(i):
    cur = torch.cuda.current_stream(); hp.wait_stream(cur)
    with torch.cuda.stream(hp): pipe.step(sets[i], outs[i])
    cur.wait_stream(hp)
res["both_train_high_prio"] = timeit(capture(step_hp))
pipe.side = torch.cuda.Stream(device=dev, priority=-1)
res["both_infer_high_prio"] = timeit(capture(lambda i: pipe.step(sets[i], outs[i])))
pipe.side = torch.cuda.Stream(device=dev)
res["both_serial"] = timeit(capture(lambda i: (pipe.infer_chain(sets[i], outs[i]), pipe.train_chain(sets[i], outs[i]))))
res["select"] = timeit(capture(k_select))
res["select+resolve"] = timeit(capture(lambda i: (k_select(i), k_resolve(i))))
res["select+resolve+pos"] = timeit(capture(lambda i: (k_select(i), k_resolve(i), k_pos(i))))
# decode alone: re-zero the counters every launch (left to accumulate, the lists would fill with DUPLICATES of the same
# locations, which k_nms' rank-by-counting does not accept: every location may be listed once per image)
res["decode"] = timeit(capture(lambda i: (pipe.cand.count.zero_(), k_decode(i))))
pipe.cand.count.zero_()
res["decode+nms"] = timeit(capture(lambda i: (k_decode(i), k_nms(i))))
for k, v in res.items():
    print(f"{k:22s} {v:8.2f} us")

# ---- cross-step overlap: independent pipelines (own scratch) on independent streams, steps alternate
def lanes_experiment(n_lanes):
    pipes = [pipe] + [DetectionHeadPipeline(levels, W, H, B, C, B * G, dev) for _ in range(n_lanes - 1)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(n_lanes)]
    graphs = []
    for ln in range(n_lanes):
        gl = []
        for i in range(3):
            o = pipes[ln].new_outputs()
            with torch.cuda.stream(streams[ln]):
                gl.append(pipes[ln].capture(sets[i], o))
        graphs.append(gl)
    torch.cuda.synchronize()
    def run(n):
        for s in range(n):
            ln = s % n_lanes
            with torch.cuda.stream(streams[ln]):
                graphs[ln][s % 3].replay()
    run(60); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cur = torch.cuda.current_stream()
    e0.record(cur)
    for st in streams: st.wait_event(e0)
    n = 900
    run(n)
    for st in streams: cur.wait_stream(st)
    e1.record(cur); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for nl in (1, 2, 3, 4):
    print(f"lanes={nl}: {lanes_experiment(nl):8.2f} us/step")

"""Developer probe: decode + NMS timings for the inference sweep (BASELINE.json configs[4]: 640-1280 px, batch 1-256,
score threshold 0.05), both decode variants, next to the reference's operator sequence as torch eager on the same GPU.
    python tools/postprocess_times.py > gpurun_out/sweep.json"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import torch_restatement as tr
from sihl_b200 import ops, synth
dev = torch.device("cuda", 0)
C, K = 80, 100
rows = []
CASES = ((640, 64, -5.0, 1.0), (640, 1, -4.0, 2.0), (640, 64, -4.0, 2.0), (640, 256, -4.0, 2.0), (896, 16, -4.0, 2.0),
         (1024, 16, -4.0, 2.0), (1280, 1, -4.0, 2.0), (1280, 16, -4.0, 2.0), (1280, 64, -4.0, 2.0), (640, 64, 2.0, 2.0))
for size, batch, mean, std in CASES:
    levels = synth.level_sizes(size, size)
    off, sc, an = ops.anchor_tables(levels, size, size, dev)
    A = an.shape[0]
    gen = torch.Generator(device=dev); gen.manual_seed(1)
    loc, iou, box, cls = synth.dense_maps_torch(gen, batch, A, C, dev, mean, std)
    cand = ops.CandidateBuffers.allocate(batch, A, dev)
    out = None
    def timeit(fn, n=20):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e3
    row = {"size": size, "batch": batch, "anchors": A, "loc_mean": mean, "loc_std": std}
    for mode in ops.DECODE_MODES:
        def decode():
            ops.dense_decode(loc, cls, box, off, sc, size, size, 0.05, cand, zero_counts=False, mode=mode)
        def nms():
            return ops.nms_topk(cand, batch, 0.5, K, out, reset_counts=True)
        cand.count.zero_(); decode(); row["candidates_per_image"] = cand.count.float().mean().item(); out = nms(); torch.cuda.synchronize()
        t_both = timeit(lambda: (decode(), nms()))
        cand.count.zero_()
        t_dec = timeit(lambda: (decode(), cand.count.zero_()))
        cand.count.zero_()
        t_split = timeit(lambda: (decode(), ops.nms_topk(cand, batch, 0.5, K, out, reset_counts=True, split=True)))
        cand.count.zero_()
        best = min(t_both, t_split)
        row[mode] = {"decode_nms_us": t_both, "decode_splitnms_us": t_split, "decode_us": t_dec, "images_per_s": batch / best * 1e6}
    if batch <= 16:
        with torch.no_grad():
            tr.dense_postprocess(levels, size, size, loc[:1], box[:1], cls[:1], 0.05, 0.5, K); torch.cuda.synchronize()
            t0 = time.perf_counter(); tr.dense_postprocess(levels, size, size, loc, box, cls, 0.05, 0.5, K); torch.cuda.synchronize()
            row["gpu_eager_reference_us"] = (time.perf_counter() - t0) * 1e6
    rows.append(row)
    print(json.dumps(row), file=sys.stderr)
print(json.dumps({"workload": "configs[4]: decode + class-aware NMS, C=80, K=100, score_thr 0.05, iou_thr 0.5", "rows": rows}, indent=1))

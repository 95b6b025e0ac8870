"""Developer probe: N1 (QuadrilateralDetection.bbox_matching for a batch) — kernels vs the reference's operator sequence
as torch eager on the same GPU and on the host CPU.  python tools/bench_quad.py > gpurun_out/quad.json"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import torch_restatement as tr
from sihl_b200 import ops, synth
from sihl_b200.heads import quadrilateral_detection as qd

dev = torch.device("cuda", 0)
H = W = 640; B, G = 64, 100
levels = synth.level_sizes(H, W)
anchors = qd.quad_anchors(levels, range(3, 8), 7, W, H, dev)
gt_np = synth.gt_batch_np(4321, B, H, W, 80, G, ragged=False)
boxes = [torch.from_numpy(b).to(dev).reshape(-1, 4) for b, _ in gt_np.per_image()]
gt = ops.GtBatch.from_lists(boxes, None, dev)
for _ in range(5): out = ops.quad_bbox_matching(anchors, gt, 9)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 50
e0.record()
for _ in range(n): out = ops.quad_bbox_matching(anchors, gt, 9)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
def eager(device, bx, an):
    t0 = time.perf_counter()
    res = [tr.quad_match_one(an, b, 9) for b in bx]
    if device != "cpu": torch.cuda.synchronize()
    return time.perf_counter() - t0, res
eager(dev, boxes[:4], anchors)
t_gpu, res = eager(dev, boxes, anchors)
same = all(torch.equal(out["o2o_mask"][b], res[b][1]) and torch.equal(out["iou"][b], res[b][2] + 0.0) and torch.equal(out["rel_iou"][b], res[b][3] + 0.0) for b in range(B))
torch.set_num_threads(os.cpu_count())
cb, ca = [b.cpu() for b in boxes[:8]], anchors.cpu()
eager("cpu", cb[:1], ca)
t_cpu, _ = eager("cpu", cb, ca)
print(json.dumps({"workload": "QuadrilateralDetection.bbox_matching, 640x640, A=8525, B=64, 100 gt/image, topk 9",
                  "ours_ms_per_batch": ms, "ours_images_per_s": B / ms * 1e3, "pairs_per_batch": B * G * anchors.shape[0],
                  "gpu_eager_reference_ms_per_batch": t_gpu * 1e3, "gpu_eager_reference_images_per_s": B / t_gpu,
                  "cpu_reference_images_per_s": 8 / t_cpu, "cpu_threads": os.cpu_count(), "bit_equal_to_gpu_eager": bool(same)}))

"""Developer probe: the training tail of the head WITH gradients at cfg1 (B=64, A=8525, C=80, 100 gt/image) —
assignment + the four losses + their backward through the drop-in head's autograd Function (sihl_b200 kernels), next to
the reference's operator sequence with torch autograd on the same GPU.  MLP outputs are synthetic leaf tensors, so only
the dense tail is timed.   python tools/bench_train_tail.py > gpurun_out/train_tail.json"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import torch_restatement as tr
from sihl_b200 import ops, synth
from sihl_b200.heads.object_detection import _DetectionLoss

dev = torch.device("cuda", 0)
H = W = 640; B, C, G = 64, 80, 100
levels = synth.level_sizes(H, W)
gen = torch.Generator(device=dev); gen.manual_seed(1234)
boxes_cat, classes_cat, offsets_t = synth.gt_batch_torch(gen, B, H, W, C, G, dev)
A = synth.num_anchors(levels)
loc, iou, box, cls = synth.dense_maps_torch(gen, B, A, C, dev)
boxes = [boxes_cat[b * G:(b + 1) * G] for b in range(B)]
classes = [classes_cat[b * G:(b + 1) * G] for b in range(B)]

def ours():
    offsets, scales, anchors = ops.anchor_tables(levels, W, H, dev)
    gt = ops.GtBatch(boxes_cat, classes_cat, offsets_t, [G] * B)
    sel = ops.assign_select(anchors, levels, W, H, gt, 9, terms=ops.anchor_terms(levels, W, H, dev))
    res = ops.assign_resolve(sel, gt, A, 9, True, want_positives=True)
    pos_index, pos_total, _ = ops.pos_compact(res["tile_pos_count"], res["tile_pos_rows"], B, A)
    P = int(pos_total.item())                                  # the step's one host sync
    pos_index = pos_index[:P]
    l = loc.detach().requires_grad_(True); i = iou.detach().requires_grad_(True)
    rows = pos_index.long()
    bx = box.view(-1, 4)[rows].detach().requires_grad_(True); cl = cls.view(-1, C)[rows].detach().requires_grad_(True)
    state = dict(rel_iou=res["iou"], assignment=res["assignment"], pos_index=pos_index, P=P, A=A, offsets=offsets, scales=scales,
                 img_w=W, img_h=H, gt=gt, reduce_sums=None)
    out = _DetectionLoss.apply(l, i, bx, cl, state)
    out[4].backward()
    return out, (l.grad, i.grad, bx.grad, cl.grad)

def eager():
    l = loc.detach().requires_grad_(True); i = iou.detach().requires_grad_(True)
    bx = box.detach().requires_grad_(True); cl = cls.detach().requires_grad_(True)
    loss, metrics, _, _ = tr.train_losses(levels, W, H, boxes, classes, l, i, bx, cl, 9)
    loss.backward()
    return loss, (l.grad, i.grad, bx.grad, cl.grad)

def timeit(fn, reps):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): r = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps, r

t_ours, (out, g_ours) = timeit(ours, 20)
t_eager, (loss, g_eager) = timeit(eager, 1)
rel = abs(float(out[4]) - float(loss)) / abs(float(loss))
gl = float((g_ours[0] - g_eager[0]).abs().max() / g_eager[0].abs().max())
print(json.dumps({"workload": "cfg1 training tail with gradients (assign + 4 losses + backward), B=64", "ours_ms": t_ours * 1e3,
                  "ours_images_per_s": B / t_ours, "gpu_eager_reference_ms": t_eager * 1e3, "gpu_eager_images_per_s": B / t_eager,
                  "loss_rel_diff": rel, "dloc_max_rel_diff": gl}))

"""Developer probe: the training tail of the head WITH gradients at cfg1 (B=64, A=8525, C=80, 100 gt/image) —
assignment + positive compaction + the four losses + their backward through the drop-in head's own code path
(``ops.train_assign`` -> ``_TrainLoss`` autograd Function -> ``backward``: 3 C calls, 5 kernel launches, no host sync),
next to the reference's operator sequence with torch autograd on the same GPU.  The MLP outputs are synthetic leaf
tensors (gathered once, outside the timed region, from dense maps at the positive rows), so only the dense tail is timed.

Reported: ``ours_ms`` eager (host-launched every step), ``ours_graph_ms`` the same step captured once into a CUDA graph
and replayed (what "graph capturable" buys), per-kernel device time is in profiles/.
    python tools/bench_train_tail.py > gpurun_out/train_tail.json"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sihl_b200 import ops, synth
from sihl_b200.heads.object_detection import _TrainLoss

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
H = W = 640; B, C, G = 64, 80, 100
dtype = {"f32": torch.float32, "bf16": torch.bfloat16, "f16": torch.float16}[os.environ.get("TAIL_DTYPE", "f32")]
levels = synth.level_sizes(H, W)
gen = torch.Generator(device=dev); gen.manual_seed(1234)
boxes_cat, classes_cat, offsets_t = synth.gt_batch_torch(gen, B, H, W, C, G, dev)
A = synth.num_anchors(levels)
loc, iou, box, cls = synth.dense_maps_torch(gen, B, A, C, dev)
loc, iou, box, cls = (t.to(dtype) for t in (loc, iou, box, cls))
box2, cls2 = box.view(-1, 4), cls.view(-1, C)
counts = [G] * B
boxes = [boxes_cat[b * G:(b + 1) * G] for b in range(B)]
classes = [classes_cat[b * G:(b + 1) * G] for b in range(B)]


# stand-ins for box_head / cls_head(o2m_feats): the rows the MLPs would produce for this (fixed) ground truth, prepared once
_st0 = ops.train_assign(levels, W, H, boxes_cat, classes_cat, counts, B, 9)
bx0 = box2.index_select(0, _st0.pos_index).contiguous()
cl0 = cls2.index_select(0, _st0.pos_index).contiguous()


def ours(gt_offsets=None):
    st = ops.train_assign(levels, W, H, boxes_cat, classes_cat, None if gt_offsets is not None else counts, B, 9,
                          gt_offsets=gt_offsets)
    l = loc.detach().requires_grad_(True); i = iou.detach().requires_grad_(True)
    bx = bx0.detach().requires_grad_(True); cl = cl0.detach().requires_grad_(True)
    out = _TrainLoss.apply(l, i, bx, cl, st, None)
    out[4].backward()
    return out, (l.grad, i.grad, bx.grad, cl.grad), st


def eager():
    from oracle import torch_restatement as tr
    l, i, bx, cl = (t.detach().float().clone().requires_grad_(True) for t in (loc, iou, box, cls))   # fresh leaves every call
    loss, metrics, _, _ = tr.train_losses(levels, W, H, boxes, classes, l, i, bx, cl, 9)
    loss.backward()
    return loss, (l.grad, i.grad, bx.grad, cl.grad)


def timeit(fn, reps):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): r = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps, r


t_ours, (out, g_ours, st) = timeit(ours, 200)
P = int(st.pos_total.item())

# the same step as ONE CUDA graph (static shapes; gt offsets on the device so a replay could carry new ground truth)
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(3):
        ours(offsets_t)
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    g_out, g_grads, g_st = ours(offsets_t)
torch.cuda.synchronize()
for _ in range(20): graph.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 500
e0.record()
for _ in range(reps): graph.replay()
e1.record(); torch.cuda.synchronize()
t_graph = e0.elapsed_time(e1) / reps * 1e-3
graph_same = bool(torch.equal(g_out, out) and all(torch.equal(a, b) for a, b in zip(g_grads, g_ours)))

t_eager, (loss, g_eager) = timeit(eager, 1)
rel = abs(float(out[4].detach()) - float(loss.detach())) / abs(float(loss.detach()))
rows = st.pos_index[:P].long()
gl = float((g_ours[0].float() - g_eager[0]).abs().max() / g_eager[0].abs().max())
gc = float((g_ours[3][:P].float() - g_eager[3].view(-1, C)[rows]).abs().max() / g_eager[3].abs().max())
gb = float((g_ours[2][:P].float() - g_eager[2].view(-1, 4)[rows]).abs().max() / g_eager[2].abs().max())
pad_zero = bool((g_ours[3][P:] == 0).all() and (g_ours[2][P:] == 0).all())
print(json.dumps({"workload": "cfg1 training tail with gradients (assign + compaction + 4 losses + backward), B=64",
                  "map_dtype": str(dtype), "positives": P, "row_capacity": st.capacity,
                  "ours_ms": t_ours * 1e3, "ours_images_per_s": B / t_ours,
                  "ours_graph_ms": t_graph * 1e3, "ours_graph_images_per_s": B / t_graph, "graph_equals_eager": graph_same,
                  "gpu_eager_reference_ms": t_eager * 1e3, "gpu_eager_images_per_s": B / t_eager,
                  "loss_rel_diff": rel, "dloc_max_rel_diff": gl, "dcls_max_rel_diff": gc, "dbox_max_rel_diff": gb,
                  "padding_rows_have_zero_gradient": pad_zero}))

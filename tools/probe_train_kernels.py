"""Developer probe: device time of the training-step kernels of the drop-in head, one entry point / role at a time
(each variant captured 20x into a CUDA graph and replayed, so host overhead is out of the numbers).
    python tools/probe_train_kernels.py [f32|bf16|f16] > gpurun_out/train_kernels.json"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sihl_b200 import ops, synth

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
dtype = {"f32": torch.float32, "bf16": torch.bfloat16, "f16": torch.float16}[sys.argv[1] if len(sys.argv) > 1 else "f32"]
H = W = 640; B, C, G = 64, 80, 100
levels = synth.level_sizes(H, W)
gen = torch.Generator(device=dev); gen.manual_seed(1234)
boxes_cat, classes_cat, offsets_t = synth.gt_batch_torch(gen, B, H, W, C, G, dev)
A = synth.num_anchors(levels)
loc, iou, box, cls = (t.to(dtype) for t in synth.dense_maps_torch(gen, B, A, C, dev))
st = ops.train_assign(levels, W, H, boxes_cat, classes_cat, [G] * B, B, 9)
bx = box.view(-1, 4).index_select(0, st.pos_index).contiguous()
cl = cls.view(-1, C).index_select(0, st.pos_index).contiguous()
torch.cuda.synchronize()


def graph_time(fn, inner=20, reps=30):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(inner):
            fn()
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps / inner * 1e3


def zero_then(fn):
    def run():
        st.sums.zero_()
        fn()
    return run

res = {"map_dtype": str(dtype), "positives": int(st.pos_total.item()), "capacity": st.capacity}
res["memset_sums_us"] = graph_time(lambda: st.sums.zero_())
res["train_assign_us (select+resolve+compact)"] = graph_time(lambda: ops.train_assign(levels, W, H, boxes_cat, classes_cat, [G] * B, B, 9))
for name, args in (("dense", (loc, iou, None, None)), ("dense+box", (loc, iou, bx, None)), ("dense+cls", (loc, iou, None, cl)),
                   ("all", (loc, iou, bx, cl))):
    res[f"train_loss[{name}]_us (incl. memset)"] = graph_time(zero_then(lambda a=args: ops.train_loss(st, *a)))
st.sums.zero_(); _, maps_all = ops.train_loss(st, loc, iou, bx, cl)
for name, want in (("dense", (True, True, False, False)), ("box", (False, False, True, False)), ("cls", (False, False, False, True)),
                   ("all", (True, True, True, True))):
    res[f"train_loss_bwd[{name}]_us"] = graph_time(lambda w=want: ops.train_loss_bwd(st, maps_all, None, w))
print(json.dumps(res, indent=1))

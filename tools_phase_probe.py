"""Developer probe: per-phase SM clocks of k_nms_small (needs the -DSIHL_PHASE_TIMING build)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
os.environ["SIHL_B200_LIB"] = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sihl_b200/lib/libsihl_b200_dbg.so")
import torch
from sihl_b200 import ops, synth, _native
from sihl_b200.pipeline import DetectionHeadPipeline, StepInputs
dev = torch.device("cuda", 0)
H = W = 640; B, Cc, G, K = 64, 80, 100, 100
levels = synth.level_sizes(H, W)
pipe = DetectionHeadPipeline(levels, W, H, B, Cc, B * G, dev)
gen = torch.Generator(device=dev); gen.manual_seed(1234)
boxes, classes, offsets = synth.gt_batch_torch(gen, B, H, W, Cc, G, dev)
loc, iou, box, cls = synth.dense_maps_torch(gen, B, pipe.A, Cc, dev)
x = StepInputs(loc, iou, box, cls, ops.GtBatch(boxes, classes, offsets, [G] * B)); out = pipe.new_outputs()
lib = _native.load()
buf = (C.c_longlong * 16)()
for it in range(5):
    pipe.infer_chain(x, out); torch.cuda.synchronize()
    lib.sihl_od_debug_phases(buf)
    t = list(buf)
    print("nms_small phases (cycles):", [t[i + 1] - t[i] for i in range(0, 6)], "tail", t[8] - t[6], t[7] - t[8], "total", t[7] - t[0])

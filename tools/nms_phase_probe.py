"""Developer probe: phase clocks of the long-list NMS body of CTA 0 (needs tools/build_dbg.sh)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["SIHL_B200_LIB"] = os.path.join(ROOT, "sihl_b200/lib/libsihl_b200_dbg.so")
import torch
from sihl_b200 import ops, synth, _native
dev = torch.device("cuda", 0)
lib = _native.load()
buf = (C.c_longlong * 16)()
for size, batch, split in ((640, 64, False), (1280, 1, False), (1280, 1, True)):
    levels = synth.level_sizes(size, size)
    off, sc, an = ops.anchor_tables(levels, size, size, dev)
    A = an.shape[0]
    gen = torch.Generator(device=dev); gen.manual_seed(1)
    loc, iou, box, cls = synth.dense_maps_torch(gen, batch, A, 80, dev, -4.0, 2.0)
    cand = ops.CandidateBuffers.allocate(batch, A, dev)
    for it in range(2):
        ops.dense_decode(loc, cls, box, off, sc, size, size, 0.05, cand, zero_counts=False)
        n0 = int(cand.count[0].item())
        ops.nms_topk(cand, batch, 0.5, 100, None, reset_counts=True, split=split); torch.cuda.synchronize()
        lib.sihl_od_debug_phases(buf)
        t = list(buf)
        print(f"S={size} split={split} n(img0)={n0}: phases", [t[i + 1] - t[i] for i in range(8)], "total", t[8] - t[0])

"""Developer probe: per-gt cycles of k_assign_select / per-phase clocks of k_nms_small (needs tools_build_dbg.sh)."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["SIHL_B200_LIB"] = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "sihl_b200/lib/libsihl_b200_dbg.so")
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
for it in range(3):
    pipe.infer_chain(x, out); torch.cuda.synchronize()
    lib.sihl_od_debug_phases(buf)
    t = list(buf)
    print("nms_small phase cycles:", [t[i + 1] - t[i] for i in range(0, 7)], "total", t[7] - t[0])
n = B * G
arr = (C.c_uint * (3 * n))()
for it in range(3):
    pipe.train_chain(x, out); torch.cuda.synchronize()
    lib.sihl_od_debug_select(arr, n)
    a = np.frombuffer(arr, dtype=np.uint32).reshape(n, 3).astype(np.int64)
    start = (a[:, 0] - a[:, 0].min()) & 0xffffffff
    cyc, chunks = a[:, 1], a[:, 2]
    end_ns = start + cyc / 1.9
    print(f"select: warps={n} cycles mean={cyc.mean():.0f} p50={np.median(cyc):.0f} p90={np.percentile(cyc,90):.0f} max={cyc.max()} | chunks mean={chunks.mean():.1f} max={chunks.max()} "
          f"| start ns p50={np.median(start):.0f} max={start.max()} | last end ns={end_ns.max():.0f} | cycles/chunk={cyc.sum()/chunks.sum():.0f}")
    big = np.argsort(-cyc)[:5]
    print("   slowest:", [(int(cyc[i]), int(chunks[i]), int(start[i])) for i in big])

nb = 148 * 12
bb = (C.c_ulonglong * (4 * nb))()
for it in range(3):
    pipe.train_chain(x, out); torch.cuda.synchronize()
    lib.sihl_od_debug_pos_blocks(bb, nb)
    a = np.frombuffer(bb, dtype=np.uint64).reshape(nb, 4).astype(np.int64)
    t0 = a[:, 0].min()
    a = a - t0
    print(f"pos blocks: start p50={np.median(a[:,0]):.0f} max={a[:,0].max()} | loop end p50={np.median(a[:,1]):.0f} max={a[:,1].max()} | accum end max={a[:,2].max()} | exit max={a[:,3].max()} ns"
          f" | loop dur mean={(a[:,1]-a[:,0]).mean():.0f} max={(a[:,1]-a[:,0]).max()}")

rb = (C.c_longlong * 128)()
for it in range(2):
    pipe.train_chain(x, out); torch.cuda.synchronize()
    lib.sihl_od_debug_resolve(rb)
    t = list(rb)
    for blk in range(8):
        q = t[blk * 16: blk * 16 + 6]
        print(f"resolve block tile={'0' if blk < 4 else '16'} img={blk % 4}: phases", [q[i + 1] - q[i] for i in range(5)], "total", q[5] - q[0])

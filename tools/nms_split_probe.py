"""Developer probe: one 1280x1280 image (about 10k candidates), single-CTA NMS vs class-split NMS (run under ncu for per-kernel times)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sihl_b200 import ops, synth
dev = torch.device("cuda", 0)
size, batch = 1280, int(sys.argv[1]) if len(sys.argv) > 1 else 1
levels = synth.level_sizes(size, size)
off, sc, an = ops.anchor_tables(levels, size, size, dev)
gen = torch.Generator(device=dev); gen.manual_seed(1)
loc, iou, box, cls = synth.dense_maps_torch(gen, batch, an.shape[0], 80, dev, -4.0, 2.0)
cand = ops.CandidateBuffers.allocate(batch, an.shape[0], dev)
for split in (False, True, False, True):
    ops.dense_decode(loc, cls, box, off, sc, size, size, 0.05, cand, zero_counts=False, mode="candidate_first")
    ops.nms_topk(cand, batch, 0.5, 100, None, reset_counts=True, split=split)
torch.cuda.synchronize()
print("ok")

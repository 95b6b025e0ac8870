"""Developer probe: dense decode + NMS timings for the inference sweep (configs[4])."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sihl_b200 import ops, synth
dev = torch.device("cuda", 0)
C, K = 80, 100
for size, batch, mean, std in ((640, 64, -5.0, 1.0), (640, 64, -4.0, 2.0), (1024, 16, -4.0, 2.0), (1280, 16, -4.0, 2.0), (1280, 1, -4.0, 2.0)):
    levels = synth.level_sizes(size, size)
    off, sc, an = ops.anchor_tables(levels, size, size, dev)
    A = an.shape[0]
    gen = torch.Generator(device=dev); gen.manual_seed(1)
    loc, iou, box, cls = synth.dense_maps_torch(gen, batch, A, C, dev, mean, std)
    cand = ops.CandidateBuffers.allocate(batch, A, dev)
    out = None
    def decode():
        ops.dense_decode(loc, cls, box, off, sc, size, size, 0.05, cand, zero_counts=False)
    def nms():
        return ops.nms_topk(cand, batch, 0.5, K, out, reset_counts=True)
    decode(); ncand = cand.count.float().mean().item(); out = nms(); torch.cuda.synchronize()
    def timeit(fn, n=30):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e3
    t_both = timeit(lambda: (decode(), nms()))
    cand.count.zero_()
    t_dec = timeit(lambda: (decode(), cand.count.zero_()))
    print(f"S={size} B={batch} A={A} cand/img={ncand:.0f}: decode+nms {t_both:8.1f} us, decode(+memset) {t_dec:8.1f} us, nms ~{t_both - t_dec:8.1f} us, {batch / t_both * 1e6:10.0f} img/s")

"""Developer probe: batched_nms (torchvision signature) timings, one image, 80 classes, vs torchvision on the same GPU."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torchvision
from sihl_b200 import ops, synth
dev = torch.device("cuda", 0)
rows = []
for n, size, ncls in ((1000, 640, 80), (2500, 640, 80), (4000, 640, 80), (5000, 1024, 80), (10000, 1024, 80), (30000, 1024, 80), (30000, 1024, 1)):
    b, s, c = synth.nms_candidates_np(7, n, size, ncls)
    b, s, c = (torch.from_numpy(x).to(dev) for x in (b, s, c))
    seg = torch.tensor([0, n], dtype=torch.int32, device=dev)
    def ours(): return ops.batched_nms(b, s, c, 0.5, seg)
    def tv(): return torchvision.ops.batched_nms(b, s, c, 0.5)
    def timeit(fn, reps=10):
        for _ in range(2): fn()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(reps): fn()
        torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e6
    k, cnt = ours()
    kept = int(cnt[0].item())
    want = tv()
    same = kept == want.numel() and bool(torch.equal(torch.sort(k[:kept]).values, torch.sort(want).values))
    rows.append({"n": n, "classes": ncls, "kept": kept, "ours_us": timeit(ours), "torchvision_cuda_us": timeit(tv), "same_keep_set": same})
    print(rows[-1], file=sys.stderr)
print(json.dumps(rows))

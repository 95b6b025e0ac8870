"""Developer probe: is the head's training step (tcgen05+train) bound by the host or by the device?  Enqueue time of N steps
(no synchronisation) against the time until the GPU has finished them, and the top host functions by cProfile."""
import cProfile, io, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sihl_b200 import synth
from sihl_b200.heads import ObjectDetection
dev = torch.device("cuda", 0)
B, size, C = 64, 640, 80
model = ObjectDetection(in_channels=[3, 64, 128, 256, 256, 256], num_classes=C, num_channels=256, num_layers=4).to(dev).train()
model.mlp_backend = sys.argv[1] if len(sys.argv) > 1 else "tcgen05+train"
g = torch.Generator(device=dev); g.manual_seed(0)
inputs = [torch.randn((B, c, max(1, size // 2 ** l), max(1, size // 2 ** l)), generator=g, device=dev) if l >= 3 or l == 0 else torch.empty((B, c, 1, 1), device=dev)
          for l, c in enumerate(model.in_channels)]
gt = synth.gt_batch_np(3, B, size, size, C, 100)
tb = [torch.from_numpy(b_).to(dev) for b_, _ in gt.per_image()]
tc = [torch.from_numpy(c_).to(dev) for _, c_ in gt.per_image()]
def step():
    model.zero_grad(set_to_none=True)
    loss, _ = model.training_step(inputs, classes=tc, boxes=tb)
    loss.backward()
for _ in range(4): step()
torch.cuda.synchronize()
N = 10
t0 = time.perf_counter()
for _ in range(N): step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"enqueue {1e3 * (t1 - t0) / N:.2f} ms/step   until done {1e3 * (t2 - t0) / N:.2f} ms/step")
pr = cProfile.Profile(); pr.enable()
for _ in range(N): step()
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(30); print(s.getvalue())

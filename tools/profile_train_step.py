"""Developer probe: kernel-time breakdown of the head's training step + backward at cfg1 (torch profiler)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
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
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:28]
tot = sum(e.device_time_total for e in prof.key_averages() if e.device_type == torch.autograd.DeviceType.CUDA) / 3e3
print(f"total device ms per step ~ {tot:.2f}")
for e in rows:
    if e.device_time_total > 0:
        print(f"{e.device_time_total / 3e3:8.3f} ms  x{e.count // 3:4d}  {e.key[:110]}")

# second view: ATen ops by input shape (SELF device time only, so nested ops are not counted twice): the torch "glue"
# between the head's own kernels
print("\n# self device time by (op, input shapes)")
rows = sorted(prof.key_averages(group_by_input_shape=True), key=lambda e: -e.self_device_time_total)[:45]
for e in rows:
    if e.self_device_time_total > 0:
        print(f"{e.self_device_time_total / 3e3:8.3f} ms  x{e.count // 3:4d}  {e.key[:60]:60s} {str(e.input_shapes)[:110]}")

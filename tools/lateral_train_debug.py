"""Developer probe: training-mode laterals on the tensor-core path vs torch, per level."""
import copy, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from sihl_b200.heads import ObjectDetection
DEV = "cuda:0"
def cos(a, b): return float(F.cosine_similarity(a.float().flatten(), b.float().flatten(), dim=0))
def rel(a, b): return float((a.float() - b.float()).norm() / b.float().norm())
torch.manual_seed(2)
size = int(sys.argv[1]) if len(sys.argv) > 1 else 96
model = ObjectDetection(in_channels=[3, 16, 32, 256, 256, 256], num_classes=5, num_channels=256, num_layers=1).to(DEV).train()
ref = copy.deepcopy(model)
g = torch.Generator().manual_seed(6)
base = [torch.randn((3, c, max(1, size // 2 ** l), max(1, size // 2 ** l)), generator=g).to(DEV) * 1.5 + 0.3 for l, c in enumerate(model.in_channels)]
xa = [t.clone().requires_grad_(True) for t in base]; xb = [t.clone().requires_grad_(True) for t in base]
model.mlp_backend = "tcgen05+train"
torch.backends.cudnn.allow_tf32 = False
flat = model._flat_feats_training(xa); want = ref._flat_feats(xb)
gg = torch.Generator(device=DEV).manual_seed(51)
gout = torch.randn(tuple(want.shape), generator=gg, device=DEV)
flat.backward(gout.bfloat16()); want.backward(gout)
off = 0
for lvl in (3, 4, 5):
    hw = xa[lvl].shape[2] * xa[lvl].shape[3]
    print("level", lvl, "rows", 3 * hw, "out rel", rel(flat[:, off:off + hw], want[:, off:off + hw]), "dx cos", cos(xa[lvl].grad, xb[lvl].grad), "rel", rel(xa[lvl].grad, xb[lvl].grad))
    off += hw
for (na, pa), (nb, pb) in zip(model.laterals.named_parameters(), ref.laterals.named_parameters()):
    print(na, "cos", cos(pa.grad, pb.grad), "rel", rel(pa.grad, pb.grad))
for la, lb in zip(model.laterals, ref.laterals):
    print("running", rel(la[1].running_mean, lb[1].running_mean), rel(la[1].running_var, lb[1].running_var))

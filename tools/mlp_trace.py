"""Developer probe: per-tile timeline of CTA 0 of the hidden-layer kernel (needs a library built with
SIHL_B200_NVCC_EXTRA=-DSIHL_MLP_TRACE).  Prints, per tile, cycles relative to the tile's first event."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sihl_b200 import ops, _native
dev = torch.device("cuda", 0)
M = 64 * 8525
x = torch.randn((M, 256), device=dev).bfloat16(); y = torch.empty_like(x)
w = (torch.randn((256, 256), device=dev) / 16).bfloat16()
b = torch.zeros(256, device=dev); g = torch.ones(256, device=dev)
for _ in range(5): ops.mlp_hidden(x, w, b, g, b, out=y)
torch.cuda.synchronize()
lib = _native.load()
fn = lib.sihl_od_mlp_debug_trace; fn.restype = ctypes.c_int; fn.argtypes = [ctypes.c_void_p, ctypes.c_int]
buf = np.zeros(64 * 16, dtype=np.int64)
n = fn(buf.ctypes.data, buf.size)
tr = buf.reshape(64, 16)
tiles = (M // 128 + 147) // 148
t0 = tr[0, 0]
names = ["req0", "req1", "req2", "req3", "accfree", "land0", "land1", "land2", "land3", "accfull", "inregs", "sum", "ssq", "done", "stored"]
print("tile " + " ".join(f"{n:>8s}" for n in names))
for t in range(min(tiles, 30) if not os.environ.get("SUMMARY") else 0):
    print(f"{t:4d} " + " ".join(f"{int(tr[t, e] - t0):8d}" for e in range(15)))
ss = tr[5:tiles - 1]
print("phases (cycles, median over steady-state tiles): wait->inregs", np.median(ss[:, 10] - ss[:, 9]), " stats+exchange", np.median(ss[:, 11] - ss[:, 10]),
      " normalise+store", np.median(ss[:, 14] - ss[:, 11]), " mma issue span", np.median(ss[:, 8] - ss[:, 4]))
per = np.diff(tr[:tiles, 9])
print("accfull period cycles: median", np.median(per), "mean", per.mean())

"""Developer probe: where the HOST time of the eager training tail goes (cProfile over tools/bench_train_tail.ours)."""
import cProfile, io, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.argv = sys.argv[:1]
import contextlib
with contextlib.redirect_stdout(io.StringIO()):
    import bench_train_tail as t
import torch
for _ in range(50): t.ours()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(300): t.ours()
torch.cuda.synchronize()
print("plain ms/step", (time.perf_counter() - t0) / 300 * 1e3)
# pieces
def piece(fn, n=300):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print("train_assign only", piece(lambda: t.ops.train_assign(t.levels, t.W, t.H, t.boxes_cat, t.classes_cat, t.counts, t.B, 9)))
st = t.ops.train_assign(t.levels, t.W, t.H, t.boxes_cat, t.classes_cat, t.counts, t.B, 9)
def fwd_only():
    with torch.no_grad():
        return t._TrainLoss.apply(t.loc, t.iou, t.bx0, t.cl0, st, None)
print("loss fwd only (no grad)", piece(fwd_only))
def leaves():
    return [x.detach().requires_grad_(True) for x in (t.loc, t.iou, t.bx0, t.cl0)]
print("4 leaves", piece(leaves))
def fwd_bwd():
    l, i, b, c = leaves()
    out = t._TrainLoss.apply(l, i, b, c, st, None)
    out[4].backward()
print("leaves + loss fwd + bwd", piece(fwd_bwd))
# the floor torch itself sets: forward + backward() of a trivial custom Function on four CUDA leaves of the same shapes
# (no kernels of ours at all): what one `loss.backward()` call costs before it reaches anybody's backward
class _Nop(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, c, d):
        ctx.save_for_backward(a, b, c, d)
        return a.new_zeros(5)
    @staticmethod
    def backward(ctx, g):
        return ctx.saved_tensors
def nop_fwd_bwd():
    l, i, b, c = leaves()
    _Nop.apply(l, i, b, c)[4].backward()
print("torch floor: 4 leaves + no-op Function + backward()", piece(nop_fwd_bwd))
pr = cProfile.Profile()
pr.enable()
for _ in range(300): t.ours()
torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(35)
print(s.getvalue())

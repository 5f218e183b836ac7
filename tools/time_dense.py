"""NumericalMarkowitz at the headline size on DIVERSIFIED problems (small expected returns: every asset ends up free), the
regime of ThorpNet / BachelierNet in training -- these go through the dense kernels, not the sparse-support ones.
    python tools/time_dense.py [rets_scale]"""
import sys
import torch
sys.path.insert(0, ".")
from deepdow_b200 import NumericalMarkowitz
dev = "cuda:0"
B, N, u = 4096, 100, 0.2
rs = float(sys.argv[1]) if len(sys.argv) > 1 else 0.02
gen = torch.Generator(device=dev).manual_seed(7)
R = torch.randn(B, N, N, generator=gen, device=dev)
C = (R @ R.transpose(1, 2) / N + 0.1 * torch.eye(N, device=dev)).contiguous()
rets = torch.randn(B, N, generator=gen, device=dev) * rs
ones = torch.ones(B, device=dev)
layer = NumericalMarkowitz(N, u)
ins = [t.requires_grad_() for t in (rets, C, ones.clone(), ones.clone())]
G = torch.randn(B, N, generator=gen, device=dev)
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
with torch.no_grad():
    tf = timeit(lambda: layer(*ins))
    w = layer(*ins)
tfb = timeit(lambda: torch.autograd.grad(layer(*ins), ins, G))
free = ((w > 0) & (w < u)).sum(1).float()
st = layer.last_status
print("solver counters (interior-point fallbacks, hand-overs to the dense kernel, unsolved):", layer.solver_counters(),
      "| status bits: fallback", int((st & 1).ne(0).sum()), "regularized", int((st & 2).ne(0).sum()), "inexact", int((st & 4).ne(0).sum()))
print(f"rets scale {rs}: mean free {free.mean().item():.1f} (max {free.max().item():.0f}); fwd {tf:.3f} ms = {B/tf/1e3:.2f} M solves/s, "
      f"fwd+bwd {tfb:.3f} ms = {B/tfb/1e3:.2f} M solves/s")

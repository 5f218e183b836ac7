"""HBM fraction of the row allocators (sparsemax / variational softmax) and of CovarianceMatrix(sqrt=False) at sizes large
enough to be bandwidth bound (run on the GPU box).  Algorithmic bytes as in SURVEY.md section 8(d)."""
import json, os, sys
import torch
sys.path.insert(0, ".")
from deepdow_b200 import CovarianceMatrix, SoftmaxAllocator, SparsemaxAllocator
dev = "cuda:0"
peak = 6548.2
pp = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(pp):
    peak = float(json.load(open(pp))["hbm_gbs"])
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
B, N = 1 << 20, 100
gen = torch.Generator(device=dev).manual_seed(3)
x = ((torch.rand(B, N, generator=gen, device=dev) - 0.5) * 10).requires_grad_()
G = torch.randn(B, N, generator=gen, device=dev)
for name, layer in (("sparsemax", SparsemaxAllocator(N, 1, 0.05)), ("softmax(variational)", SoftmaxAllocator(1, "variational", N, 0.05))):
    with torch.no_grad():
        tf = timeit(lambda: layer(x))
    w = layer(x)
    tb = timeit(lambda: torch.autograd.grad(w, x, G, retain_graph=True))
    fb, bb = (2 * N + 1) * 4 * B, 3 * N * 4 * B
    print(f"{name:22s} B={B} N={N}: fwd {tf:.3f} ms = {fb/tf/1e6:.0f} GB/s ({fb/tf/1e6/peak:.0%} of {peak:.0f}), "
          f"bwd {tb:.3f} ms = {bb/tb/1e6:.0f} GB/s ({bb/tb/1e6/peak:.0%})")
Bc, L, Nc = 1 << 15, 40, 50
xc = torch.randn(Bc, L, Nc, generator=gen, device=dev) * 0.01
cov = CovarianceMatrix(sqrt=False, shrinkage_strategy="diagonal")
with torch.no_grad():
    tc = timeit(lambda: cov(xc))
cb = (L * Nc + Nc * Nc + 1) * 4 * Bc
print(f"covariance(sqrt=False)  B={Bc} L={L} N={Nc}: fwd {tc:.3f} ms = {cb/tc/1e6:.0f} GB/s ({cb/tc/1e6/peak:.0%})")

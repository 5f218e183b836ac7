"""One small call of every kernel family (B <= 8) for compute-sanitizer:

    compute-sanitizer --tool memcheck  python tools/sanitize_smoke.py
    compute-sanitizer --tool racecheck python tools/sanitize_smoke.py

ONE tool per gpurun call (B200_PROFILING.md).  `--skip-tc` leaves the tcgen05 kernels out (a sanitizer that cannot
instrument them would otherwise hide the rest)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from deepdow_b200 import (CovarianceMatrix, NumericalMarkowitz, NumericalRiskBudgeting, SoftmaxAllocator, SparsemaxAllocator)
from deepdow_b200.benchmarks import MaximumReturn

dev = "cuda:0"
skip_tc = "--skip-tc" in sys.argv
rng = np.random.default_rng(0)


def t(a, grad=False, dtype=torch.float32):
    return torch.tensor(np.asarray(a), dtype=dtype, device=dev, requires_grad=grad)


def markowitz(B, N, u, rscale, dtype=torch.float32):
    R = rng.standard_normal((B, N, N))
    C = R @ R.transpose(0, 2, 1) / N + 0.1 * np.eye(N)
    ins = [t(rng.standard_normal((B, N)) * rscale, True, dtype), t(C, True, dtype), t(rng.uniform(0.5, 2, B), True, dtype),
           t(np.ones(B), True, dtype)]
    layer = NumericalMarkowitz(N, u)
    layer.check_status = True
    w = layer(*ins)
    torch.autograd.grad(w, ins, torch.ones_like(w))
    return w


done = []
markowitz(8, 20, 1.0, 0.5); done.append("markowitz sparse-support fwd/bwd N=20")
markowitz(8, 100, 0.2, 0.5); done.append("markowitz sparse-support fwd/bwd N=100")
markowitz(4, 100, 0.2, 0.02); done.append("markowitz dense fwd + saved-factor bwd N=100")
markowitz(4, 100, 0.2, 0.02, torch.float64); done.append("markowitz dense f64")
markowitz(2, 150, 0.05, 0.02); done.append("markowitz 256-thread dense N=150")
if not skip_tc:
    markowitz(2, 240, 0.05, 0.02); done.append("markowitz global Q (tensor-core Gram) N=240")
x = t(rng.standard_normal((4, 12, 9)) * 0.01, True)
for strategy in ("diagonal", "scaled_identity", "identity", None):
    for sq in (False, True):
        y = CovarianceMatrix(sqrt=sq, shrinkage_strategy=strategy)(x)
        y.sum().backward()
done.append("covariance plain / general / sqrt fwd+bwd N=9")
x = t(rng.standard_normal((2, 30, 140)) * 0.01, True)
CovarianceMatrix(sqrt=True)(x).sum().backward(); done.append("covariance sqrt in shared memory N=140")
if not skip_tc:
    x = t(rng.standard_normal((2, 40, 200)) * 0.01)
    CovarianceMatrix(sqrt=True)(x); CovarianceMatrix(sqrt=False, shrinkage_strategy=None)(t(rng.standard_normal((2, 40, 300)) * 0.01))
    done.append("covariance tensor-core Gram + blocked Jacobi N=200/300")
for N, u in ((5, 1.0), (100, 0.05), (300, 0.01)):
    xr = t(rng.standard_normal((8, N)), True)
    for layer in (SparsemaxAllocator(N, 1.0, u), SoftmaxAllocator(1.0, "variational", N, u)):
        layer(xr).square().sum().backward()
    MaximumReturn(max_weight=u, n_assets=N)(t(rng.standard_normal((8, 1, 4, N))))
done.append("row allocators (sparsemax / softmax / argmax) N=5,100,300")
for N, u in ((6, 1.0), (50, 0.05), (150, 0.02)):
    R = rng.random((4, N, N)); A = R @ R / N + np.eye(N)
    b = rng.random((4, N)); b /= b.sum(1, keepdims=True)
    tA, tb = t(A, True), t(b, True)
    NumericalRiskBudgeting(N, u)(tA, tb).square().sum().backward()
done.append("risk budgeting fwd/bwd N=6,50,150")
from deepdow_b200.host import MarkowitzHostPipeline
N = 20
R = rng.standard_normal((8, N, N)).astype(np.float32)
host = [torch.tensor(a_, dtype=torch.float32).pin_memory() for a_ in (rng.standard_normal((8, N)) * 0.5, R @ R.transpose(0, 2, 1) / N + 0.1 * np.eye(N),
                                                                       np.ones(8), np.ones(8), rng.standard_normal((8, N)))]
pipe = MarkowitzHostPipeline(N, 0.3, chunk=4, device=dev)
pipe.step(*host); pipe.close(); done.append("host-buffer pipeline")
torch.cuda.synchronize()
print("sanitize_smoke ok:", "; ".join(done))

"""One CovarianceMatrix(sqrt=True) forward at a large universe (for ncu launch lists).  Usage: cov_one_large.py [B] [L] [N]"""
import sys
import torch
sys.path.insert(0, ".")
from deepdow_b200 import CovarianceMatrix
B, L, N = (int(a) for a in (sys.argv[1:4] + ["256", "250", "500"][len(sys.argv) - 1:]))
x = torch.randn(B, L, N, generator=torch.Generator(device="cuda:0").manual_seed(5), device="cuda:0") * 0.01
layer = CovarianceMatrix(sqrt=True, shrinkage_strategy="diagonal")
y = layer(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); y = layer(x); e1.record(); torch.cuda.synchronize()
print(f"B={B} L={L} N={N}: {e0.elapsed_time(e1):.2f} ms, finite {bool(torch.isfinite(y).all())}")

"""One CovarianceMatrix(sqrt=False) forward at the roofline size (for ncu captures): python tools/cov_one.py"""
import sys
import torch
sys.path.insert(0, ".")
from deepdow_b200 import CovarianceMatrix
Bc, L, Nc = 1 << 15, 40, 50
gen = torch.Generator(device="cuda:0").manual_seed(3)
xc = torch.randn(Bc, L, Nc, generator=gen, device="cuda:0") * 0.01
cov = CovarianceMatrix(sqrt=False, shrinkage_strategy="diagonal")
with torch.no_grad():
    for _ in range(2):
        cov(xc)
torch.cuda.synchronize()

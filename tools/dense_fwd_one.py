"""One dense-regime NumericalMarkowitz forward (B=4096, N=100, ~86 free) for ncu."""
import sys
import torch
sys.path.insert(0, ".")
from deepdow_b200 import NumericalMarkowitz
dev = "cuda:0"
B, N, u = 4096, 100, 0.2
gen = torch.Generator(device=dev).manual_seed(7)
R = torch.randn(B, N, N, generator=gen, device=dev)
C = (R @ R.transpose(1, 2) / N + 0.1 * torch.eye(N, device=dev)).contiguous()
rets = torch.randn(B, N, generator=gen, device=dev) * 0.02
ones = torch.ones(B, device=dev)
layer = NumericalMarkowitz(N, u)
with torch.no_grad():
    for _ in range(3):
        w = layer(rets, C, ones, ones)
torch.cuda.synchronize()
print("ok", float(w.sum()))

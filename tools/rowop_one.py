"""One forward of each row allocator at the roofline size (for ncu captures): python tools/rowop_one.py"""
import sys
import torch
sys.path.insert(0, ".")
from deepdow_b200 import SoftmaxAllocator, SparsemaxAllocator
B, N = 1 << 20, 100
gen = torch.Generator(device="cuda:0").manual_seed(3)
x = (torch.rand(B, N, generator=gen, device="cuda:0") - 0.5) * 10
with torch.no_grad():
    for _ in range(2):
        SparsemaxAllocator(N, 1, 0.05)(x)
        SoftmaxAllocator(1, "variational", N, 0.05)(x)
torch.cuda.synchronize()

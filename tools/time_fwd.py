"""Forward-only time of NumericalMarkowitz at the headline size (run on the GPU box; DDW_LIB_PATH selects the build)."""
import os, sys
import torch
sys.path.insert(0, ".")
from deepdow_b200 import NumericalMarkowitz
dev = "cuda:0"
B, N, u = 4096, 100, 0.2
gen = torch.Generator(device=dev).manual_seed(4000)
R = torch.randn(B, N, N, generator=gen, device=dev); C = (R @ R.transpose(1, 2) / N + 0.1 * torch.eye(N, device=dev)).contiguous()
rets = torch.randn(B, N, generator=gen, device=dev) * 0.5; g = torch.ones(B, device=dev); a = torch.ones(B, device=dev)
layer = NumericalMarkowitz(N, u); layer.check_status = False
with torch.no_grad():
    for _ in range(5): layer(rets, C, g, a)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): layer(rets, C, g, a)
    e1.record(); torch.cuda.synchronize()
print(f"{os.environ.get('DDW_LIB_PATH', 'default'):50s} fwd {e0.elapsed_time(e1) / 50 * 1e3:.1f} us")

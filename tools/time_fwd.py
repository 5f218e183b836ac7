"""Forward-only time of NumericalMarkowitz at the headline size (run on the GPU box; DDW_LIB_PATH selects the build)."""
import os, sys
import torch
sys.path.insert(0, ".")
from deepdow_b200 import NumericalMarkowitz
dev = "cuda:0"
B, N, u = int(os.environ.get("DDW_TIME_B", "4096")), 100, 0.2
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
    w = layer(rets, C, g, a)
Q = C.double().transpose(1, 2) @ C.double() + torch.eye(N, device=dev, dtype=torch.float64)
obj = ((w.double()[:, None, :] @ Q @ w.double()[:, :, None])[:, 0, 0] - (rets.double() * w.double()).sum(1)).sum().item()
ins = [t.clone().requires_grad_() for t in (rets, C, g, a)]
G = torch.randn(B, N, generator=gen, device=dev)
for _ in range(3):
    torch.autograd.grad(layer(*ins), ins, G)
evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(30)]
for e in evs:
    wg = layer(*ins)
    e[0].record(); grads = torch.autograd.grad(wg, ins, G); e[1].record()
torch.cuda.synchronize()
bwd_us = sum(a_.elapsed_time(b_) for a_, b_ in evs) / len(evs) * 1e3
gsum = sum(float(t.double().abs().sum()) for t in grads)
print(f"{os.path.basename(os.environ.get('DDW_LIB_PATH', 'default')):28s} fwd {e0.elapsed_time(e1) / 50 * 1e3:.1f} us   objective sum {obj:.9f}  "
      f"budget err {float((w.double().sum(1) - 1).abs().max()):.1e}  status!=0 {int(layer.last_status.ne(0).sum())}  "
      f"| bwd {bwd_us:.1f} us  |grads| {gsum:.6f}")

"""Where does the host time of one fwd+bwd call go?  (run on the GPU box)"""
import sys, time
import torch
sys.path.insert(0, ".")
from deepdow_b200 import NumericalMarkowitz, _C
dev = "cuda:0"
B, N = 4096, 100
R = torch.randn(B, N, N, device=dev); C = (R @ R.transpose(1, 2) / N + 0.1 * torch.eye(N, device=dev)).contiguous()
rets = torch.randn(B, N, device=dev) * 0.5; g = torch.ones(B, device=dev); a = torch.ones(B, device=dev); G = torch.randn(B, N, device=dev)
ins = [t.requires_grad_() for t in (rets, C, g, a)]
for mode in ("lazy", False, True):
    layer = NumericalMarkowitz(N, 0.2); layer.check_status = mode
    for _ in range(5):
        w = layer(*ins); torch.autograd.grad(w, ins, G)
    torch.cuda.synchronize()
    n = 50
    t0 = time.perf_counter()
    tf = tb = 0.0
    for _ in range(n):
        a0 = time.perf_counter(); w = layer(*ins); a1 = time.perf_counter(); torch.autograd.grad(w, ins, G); a2 = time.perf_counter()
        tf += a1 - a0; tb += a2 - a1
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"mode={mode}: host fwd {1e6*tf/n:.0f} us, host bwd {1e6*tb/n:.0f} us, enqueue total {1e6*(t1-t0)/n:.0f} us/step, wall incl. drain {1e6*(t2-t0)/n:.0f} us/step")
lib = _C.lib()
t0 = time.perf_counter()
for _ in range(1000): lib.ddw_version()
print(f"ctypes call: {1e6*(time.perf_counter()-t0)/1000:.2f} us")
t0 = time.perf_counter()
for _ in range(1000): torch.empty(4096, 100, device=dev)
print(f"torch.empty: {1e6*(time.perf_counter()-t0)/1000:.2f} us")
t0 = time.perf_counter()
for _ in range(1000): e = torch.cuda.Event(); e.record()
print(f"event create+record: {1e6*(time.perf_counter()-t0)/1000:.2f} us")

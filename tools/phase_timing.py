"""Per-phase cycle breakdown of the sparse forward kernel (debug build: DDW_PHASE_TIMING=1 python -m deepdow_b200.build --force).
Run on the GPU box.  Prints average cycles per portfolio spent by thread 0 between phase marks."""
import ctypes, sys
import torch
sys.path.insert(0, ".")
from deepdow_b200 import NumericalMarkowitz, _C
dev = "cuda:0"
B, N, u = 4096, 100, 0.2
gen = torch.Generator(device=dev).manual_seed(4000)
R = torch.randn(B, N, N, generator=gen, device=dev); C = (R @ R.transpose(1, 2) / N + 0.1 * torch.eye(N, device=dev)).contiguous()
rets = torch.randn(B, N, generator=gen, device=dev) * 0.5; g = torch.ones(B, device=dev); a = torch.ones(B, device=dev)
layer = NumericalMarkowitz(N, u)
lib = ctypes.CDLL(_C.LIB_PATH)
buf = (ctypes.c_ulonglong * 16)()
with torch.no_grad():
    for _ in range(3): layer(rets, C, g, a)
    lib.ddw_debug_fwd_phases_f32(buf, 1)
    layer(rets, C, g, a)
    lib.ddw_debug_fwd_phases_f32(buf, 0)
names = ["load(TMA)", "gamma+diag", "newton init", "lists", "gram+tU+slots", "gather", "kkt solve", "t=Aw", "colpass+update", "-"]
n = max(buf[11], 1)
print(f"problems {buf[11]}, iterations/problem {buf[10]/n:.2f}, mean free set per iteration {buf[12]/max(buf[10],1):.1f}, "
      f"first-iteration free set {buf[14]/n:.1f}, newton iterations/problem {buf[9]/n:.1f}, "
      f"kkt: factor {buf[13]/n:.0f} + store/backsub {buf[15]/n:.0f} cycles/problem")
tot = sum(buf[i] for i in range(9))
for i in range(9):
    print(f"  {names[i]:16s} {buf[i]/n:9.0f} cycles/problem  {100*buf[i]/tot:5.1f}%")
print(f"  total            {tot/n:9.0f} cycles/problem = {tot/n/1.965e3:.1f} us at 1965 MHz")

# ---- backward
G = torch.randn(B, N, generator=gen, device=dev)
ins = [t.requires_grad_() for t in (rets, C, g, a)]
for _ in range(3):
    w = layer(*ins); torch.autograd.grad(w, ins, G)
lib.ddw_debug_bwd_phases_f32(buf, 1)
w = layer(*ins); torch.autograd.grad(w, ins, G)
lib.ddw_debug_bwd_phases_f32(buf, 0)
names = ["load(TMA)", "gamma+lists", "gram", "factor", "finish(backsub)", "d_rets+p,q", "dC write"]
n = max(buf[11], 1)
tot = sum(buf[i] for i in range(7))
print("backward:")
for i in range(7):
    print(f"  {names[i]:16s} {buf[i]/n:9.0f} cycles/problem  {100*buf[i]/tot:5.1f}%")
print(f"  total            {tot/n:9.0f} cycles/problem = {tot/n/1.965e3:.1f} us at 1965 MHz")

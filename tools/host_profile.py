"""cProfile of the host side of one NumericalMarkowitz fwd+bwd step (run on the GPU box)."""
import cProfile, pstats, sys, time
import torch
sys.path.insert(0, ".")
from deepdow_b200 import NumericalMarkowitz
dev = "cuda:0"
B, N = 4096, 100
R = torch.randn(B, N, N, device=dev); C = (R @ R.transpose(1, 2) / N + 0.1 * torch.eye(N, device=dev)).contiguous()
rets = torch.randn(B, N, device=dev) * 0.5; g = torch.ones(B, device=dev); a = torch.ones(B, device=dev); G = torch.randn(B, N, device=dev)
ins = [t.requires_grad_() for t in (rets, C, g, a)]
layer = NumericalMarkowitz(N, 0.2)
def step():
    w = layer(*ins); torch.autograd.grad(w, ins, G)
for _ in range(10): step()
torch.cuda.synchronize()
n = 300
t0 = time.perf_counter()
for _ in range(n): step()
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"enqueue {1e6*(t1-t0)/n:.0f} us/step, wall incl. drain {1e6*(t2-t0)/n:.0f} us/step")
pr = cProfile.Profile(); pr.enable()
for _ in range(n): step()
pr.disable(); torch.cuda.synchronize()
st = pstats.Stats(pr); st.sort_stats("tottime").print_stats(28)

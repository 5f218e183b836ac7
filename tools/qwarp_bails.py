"""How many portfolios of the bench workload leave the warp-per-portfolio forward for the dense kernel, per bench rank seed (GPU box)."""
import sys
import torch
sys.path.insert(0, ".")
from deepdow_b200 import NumericalMarkowitz
dev = "cuda:0"
B, N, u = 4096, 100, 0.2
for rank in range(8):
    gen = torch.Generator(device=dev).manual_seed(1000 * 4 + rank)
    R = torch.randn(B, N, N, generator=gen, device=dev)
    C = (R @ R.transpose(1, 2) / N + 0.1 * torch.eye(N, device=dev)).contiguous()
    rets = torch.randn(B, N, generator=gen, device=dev) * 0.5
    layer = NumericalMarkowitz(N, u); layer.check_status = True
    with torch.no_grad():
        w = layer(rets, C, torch.ones(B, device=dev), torch.ones(B, device=dev))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): layer(rets, C, torch.ones(B, device=dev), torch.ones(B, device=dev))
        e1.record(); torch.cuda.synchronize()
    print(f"rank seed {rank}: counters (ipm, dense hand-overs, unsolved) {layer.solver_counters()}  fwd {e0.elapsed_time(e1) / 20 * 1e3:.1f} us  max free {int((layer.last_state == 0).sum(1).max()) if hasattr(layer, 'last_state') else '?'}")

"""End-to-end step time of the host-buffer pipeline at the headline size (run on the GPU box)."""
import os, sys
import torch
sys.path.insert(0, ".")
from deepdow_b200.host import MarkowitzHostPipeline
dev = "cuda:0"
B, N, u = 4096, 100, 0.2
chunk = int(sys.argv[1]) if len(sys.argv) > 1 else 128
gen = torch.Generator().manual_seed(4000)
R = torch.randn(B, N, N, generator=gen); C = (R @ R.transpose(1, 2) / N + 0.1 * torch.eye(N)).contiguous().pin_memory()
rets = (torch.randn(B, N, generator=gen) * 0.5).pin_memory(); g = torch.ones(B).pin_memory(); a = torch.ones(B).pin_memory()
G = torch.randn(B, N, generator=gen).pin_memory()
pipe = MarkowitzHostPipeline(N, u, chunk=chunk, device=dev)
out = pipe.step(rets, C, g, a, G)
for _ in range(3): out = pipe.step(rets, C, g, a, G, out=out, sync=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(30): out = pipe.step(rets, C, g, a, G, out=out, sync=False)
e1.record(); torch.cuda.synchronize(); pipe.check()
print(f"chunk {chunk} ramp {'off' if os.environ.get('DDW_HOST_PIPELINE_NO_RAMP') else 'on'}: {e0.elapsed_time(e1)/30:.3f} ms per step")

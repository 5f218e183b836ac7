"""torch.profiler view of the headline fwd+bwd step replayed as a CUDA graph: per-kernel device time INSIDE the graph (ncu launch
lists are cold-cache and serialised), so that gaps between the kernel sum and the step time show up (GPU box)."""
import sys
import torch
from torch.profiler import profile, ProfilerActivity
sys.path.insert(0, ".")
from deepdow_b200 import NumericalMarkowitz
dev = "cuda:0"
B, N, u = 4096, 100, 0.2
gen = torch.Generator(device=dev).manual_seed(4000)
R = torch.randn(B, N, N, generator=gen, device=dev); C = (R @ R.transpose(1, 2) / N + 0.1 * torch.eye(N, device=dev)).contiguous()
rets = torch.randn(B, N, generator=gen, device=dev) * 0.5
G = torch.randn(B, N, generator=gen, device=dev)
ins = [t.requires_grad_() for t in (rets, C, torch.ones(B, device=dev), torch.ones(B, device=dev))]
layer = NumericalMarkowitz(N, u); layer.check_status = False
side = torch.cuda.Stream()
with torch.cuda.stream(side):
    for _ in range(3): torch.autograd.grad(layer(*ins), ins, G)
torch.cuda.current_stream().wait_stream(side)
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    grads = torch.autograd.grad(layer(*ins), ins, G)
for _ in range(5): graph.replay()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(20): graph.replay()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
# one replay = the events between two consecutive probe kernels
starts = [i for i, e in enumerate(ev) if "qprobe" in e.name]
i0, i1 = starts[10], starts[11]
t0 = ev[i0].time_range.start
prev_end = t0
for e in ev[i0:i1]:
    print(f"{e.time_range.start - t0:8.1f} us  +{e.time_range.start - prev_end:5.1f} gap  {e.time_range.end - e.time_range.start:7.1f} us  {e.name[:70]}")
    prev_end = e.time_range.end
print(f"step (probe to probe): {ev[i1].time_range.start - t0:.1f} us")

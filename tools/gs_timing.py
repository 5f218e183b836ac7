"""Cycles per portfolio that each role of markowitz_gram_small_kernel spends in its waits and work sections.
Build with DDW_GS_TIMING=1 python -m deepdow_b200.build (objects in build_gst/), run on the GPU box with DDW_QWARP=1."""
import ctypes, os, sys
import torch
sys.path.insert(0, ".")
from deepdow_b200 import NumericalMarkowitz, _C
dev = "cuda:0"
B, N, u = 4096, 100, 0.2
gen = torch.Generator(device=dev).manual_seed(4000)
R = torch.randn(B, N, N, generator=gen, device=dev); C = (R @ R.transpose(1, 2) / N + 0.1 * torch.eye(N, device=dev)).contiguous()
rets = torch.randn(B, N, generator=gen, device=dev) * 0.5; g = torch.ones(B, device=dev); a = torch.ones(B, device=dev)
if os.environ.get("GS_RANDOM_GAMMA"): g = 0.5 + torch.rand(B, generator=gen, device=dev)
layer = NumericalMarkowitz(N, u); layer.check_status = False
lib = _C.lib() if hasattr(_C, "lib") else _C._load()
out = (ctypes.c_ulonglong * 16)()
with torch.no_grad():
    for _ in range(3): layer(rets, C, g, a)
    lib.ddw_debug_gs_phases(out, 1)
    reps = 10
    for _ in range(reps): layer(rets, C, g, a)
    lib.ddw_debug_gs_phases(out, 0)
names = ["tma: wait raw_empty", "mma: wait acc_empty", "mma: wait ring_full", "mma: issue+commit", "stage: group sync", "stage: wait raw_full",
         "stage: gamma", "stage: wait ring_empty", "stage: split+store", "stage: fence+arrive", "stage: diag+release", "epi: wait acc_full", "epi: load+store"]
for n, v in zip(names, out):
    print(f"  {n:26s} {v / reps / B:9.0f} cycles/portfolio")

tr = (ctypes.c_longlong * 4096)()
lib.ddw_debug_gs_trace(tr)
ev = []
for role, name in enumerate(["stageA", "stageB", "mma", "epi"]):
    for it in range(4):
        for c in range(8):
            b0, e0 = tr[((role * 4 + it) * 8 + c) * 2], tr[((role * 4 + it) * 8 + c) * 2 + 1]
            if b0: ev.append((b0, e0, name, it, c))
ev.sort()
t0 = ev[0][0]
print("timeline of CTA 0 (cycles from its first event): role portfolio.chunk begin..end")
for b0, e0, name, it, c in ev:
    print(f"  {b0 - t0:8d} .. {e0 - t0:8d}  {name:7s} {it}.{c}")

print("MMA thread, portfolio 2 of CTA 0: cycles between consecutive instructions (12 tcgen05.mma, commit(stage pair), commit(accumulator))")
for c in range(4):
    ts = [tr[256 + c * 16 + i] for i in range(14)]
    begin = tr[((2 * 4 + 2) * 8 + c) * 2]
    prev = begin
    out_ = []
    for t_ in ts:
        out_.append(f"{t_ - prev:5d}" if t_ and prev else "    -")
        if t_: prev = t_
    end_prev = tr[((2 * 4 + (2 if c else 1)) * 8 + (c - 1 if c else 3)) * 2 + 1]
    print(f"  chunk {c}: " + " ".join(out_) + f"   | previous section end -> flag seen {tr[256 + c * 16 + 14] - end_prev}, -> after fence {tr[256 + c * 16 + 15] - end_prev}")
print("flag value minus own chunk index when first seen, and spins:", [(tr[320 + c], tr[328 + c]) for c in range(4)])

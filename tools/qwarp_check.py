"""Warp-per-portfolio forward (markowitz_qwarp.cu) against the float64 oracle: weights, states, solver counters (GPU box)."""
import os
import sys
os.environ.setdefault("DDW_QWARP_MIN_BATCH", "1")      # the tensor-core path also for these small batches
import numpy as np
import torch
sys.path.insert(0, ".")
import oracle
from deepdow_b200 import NumericalMarkowitz
from bench import make_inputs_numpy

dev = "cuda:0"
for (B, N, u, scale) in [(256, 100, 0.2, 1.0), (64, 20, 1.0, 1.0), (128, 50, 0.5, 0.3), (64, 128, 0.1, 1.0), (32, 37, 0.3, 1.0), (256, 100, 0.2, 0.04)]:
    rets, C, g, a, _ = make_inputs_numpy(B, N, 7 + N)
    rets = (rets * scale).astype(np.float32)
    rng = np.random.default_rng(N)
    g = (0.5 + rng.random(B)).astype(np.float32); a = (rng.random(B) * 2 - 1).astype(np.float32)
    layer = NumericalMarkowitz(N, u); layer.check_status = True
    with torch.no_grad():
        w = layer(*[torch.tensor(t, device=dev) for t in (rets, C, g, a)])
    st = layer.last_state.cpu().numpy() if hasattr(layer, "last_state") else None
    w_ref, st_ref = oracle.markowitz_fwd(rets.astype(np.float64), C.astype(np.float64), g.astype(np.float64), a.astype(np.float64), u)[:2]
    err = np.abs(w.double().cpu().numpy() - w_ref).max()
    free = (np.asarray(st_ref) == 0).sum(1)
    print(f"B={B} N={N} u={u} scale={scale}: max |w - w_ref| = {err:.2e}  free mean {free.mean():.1f} max {free.max()}  "
          f"status!=0 {int(layer.last_status.ne(0).sum())}  counters {layer.solver_counters() if hasattr(layer, 'solver_counters') else ''}")

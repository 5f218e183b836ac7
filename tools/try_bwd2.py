import sys
import numpy as np, torch
sys.path.insert(0, ".")
import oracle
from deepdow_b200 import NumericalMarkowitz
dev = "cuda:0"
rng = np.random.default_rng(0)
for (B, N, u) in [(4, 100, 1.0), (4, 128, 1.0), (4, 140, 1.0), (4, 200, 1.0), (4, 240, 1.0), (4, 256, 1.0), (4, 300, 1.0)]:
    R = rng.standard_normal((B, N, N)); C = R @ R.transpose(0, 2, 1) / N + 0.1 * np.eye(N)
    rets = rng.standard_normal((B, N)) * 0.02
    G = rng.standard_normal((B, N))
    w_ref, st, kkt, status = oracle.markowitz_fwd(rets, C, np.ones(B), np.ones(B), u)
    d_r, d_C, d_g, d_a = oracle.markowitz_bwd(G, w_ref, st, C, np.ones(B), np.ones(B), u)
    for dt in (torch.float64, torch.float32):
        tr = torch.tensor(rets, dtype=dt, device=dev, requires_grad=True)
        tC = torch.tensor(C, dtype=dt, device=dev, requires_grad=True)
        g = torch.ones(B, device=dev, dtype=dt, requires_grad=True); a = torch.ones(B, device=dev, dtype=dt, requires_grad=True)
        w = NumericalMarkowitz(N, u)(tr, tC, g, a)
        w.backward(torch.tensor(G, dtype=dt, device=dev))
        rel = lambda a_, b_: np.abs(a_.double().cpu().numpy() - b_).max() / max(np.abs(b_).max(), 1e-30)
        print(f"{str(dt)[6:]} N={N} nfree {(st==0).sum(1).mean():.1f}: werr {np.abs(w.detach().double().cpu().numpy()-w_ref).max():.1e} d_rets {rel(tr.grad,d_r):.1e} d_C {rel(tC.grad,d_C):.1e} d_g {rel(g.grad,d_g):.1e} d_a {rel(a.grad,d_a):.1e}", flush=True)

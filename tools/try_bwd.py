import sys
import numpy as np, torch
sys.path.insert(0, ".")
import oracle
from deepdow_b200 import NumericalMarkowitz
dev = "cuda:0"
rng = np.random.default_rng(0)
for (B, N, u, rep) in [(8, 240, 0.05, 2), (8, 232, 0.05, 1), (8, 256, 0.05, 1), (8, 236, 1.0, 1), (8, 200, 0.02, 1), (8, 228, 0.02, 1)]:
    R = rng.standard_normal((B, N, N)); C = R @ R.transpose(0, 2, 1) / N + 0.1 * np.eye(N)
    rets = rng.standard_normal((B, N)) * 0.5
    G = rng.standard_normal((B, N))
    w_ref, st, kkt, status = oracle.markowitz_fwd(rets, C, np.ones(B), np.ones(B), u)
    d_r, d_C, d_g, d_a = oracle.markowitz_bwd(G, w_ref, st, C, np.ones(B), np.ones(B), u)
    for r in range(rep):
        tr = torch.tensor(rets, dtype=torch.float64, device=dev, requires_grad=True)
        tC = torch.tensor(C, dtype=torch.float64, device=dev, requires_grad=True)
        g = torch.ones(B, device=dev, dtype=torch.float64, requires_grad=True); a = torch.ones(B, device=dev, dtype=torch.float64, requires_grad=True)
        w = NumericalMarkowitz(N, u)(tr, tC, g, a)
        w.backward(torch.tensor(G, dtype=torch.float64, device=dev))
        rel = lambda a_, b_: np.abs(a_.double().cpu().numpy() - b_).max() / max(np.abs(b_).max(), 1e-30)
        print(f"f64 N={N} u={u} nfree {(st==0).sum(1).mean():.1f}: werr {np.abs(w.detach().cpu().numpy()-w_ref).max():.1e} d_rets {rel(tr.grad,d_r):.1e} d_C {rel(tC.grad,d_C):.1e} d_g {rel(g.grad,d_g):.1e} d_a {rel(a.grad,d_a):.1e}", flush=True)
        tr = torch.tensor(rets, dtype=torch.float32, device=dev, requires_grad=True)
        tC = torch.tensor(C, dtype=torch.float32, device=dev, requires_grad=True)
        g = torch.ones(B, device=dev, requires_grad=True); a = torch.ones(B, device=dev, requires_grad=True)
        w = NumericalMarkowitz(N, u)(tr, tC, g, a)
        w.backward(torch.tensor(G, dtype=torch.float32, device=dev))
        print(f"f32 N={N} u={u}: werr {np.abs(w.detach().double().cpu().numpy()-w_ref).max():.1e} d_rets {rel(tr.grad,d_r):.1e} d_C {rel(tC.grad,d_C):.1e} d_g {rel(g.grad,d_g):.1e} d_a {rel(a.grad,d_a):.1e}", flush=True)

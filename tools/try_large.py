"""Ad-hoc check of the large-N (global workspace) variants against the oracle; run on the GPU box."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import oracle
from deepdow_b200 import NumericalMarkowitz, CovarianceMatrix

dev = "cuda:0"
rng = np.random.default_rng(0)
for (B, L, N, u) in [(8, 250, 240, 0.05), (8, 250, 300, 0.02)]:
    x = rng.standard_normal((B, L, N)) * 0.01
    rets = rng.standard_normal((B, N)) * 0.01
    t0 = time.time()
    tx = torch.tensor(x, dtype=torch.float32, device=dev)
    C = CovarianceMatrix(sqrt=True, shrinkage_strategy="diagonal")(tx)
    torch.cuda.synchronize(); t1 = time.time()
    ref = oracle.covariance_fwd(tx.double().cpu().numpy(), 0.5, "diagonal", True)
    err_c = np.abs(C.double().cpu().numpy() - ref).max() / np.abs(ref).max()
    tr = torch.tensor(rets, dtype=torch.float32, device=dev, requires_grad=True)
    tC = C.detach().clone().requires_grad_()
    g = torch.ones(B, device=dev, requires_grad=True); a = torch.ones(B, device=dev, requires_grad=True)
    layer = NumericalMarkowitz(N, u)
    torch.cuda.synchronize(); t2 = time.time()
    w = layer(tr, tC, g, a)
    torch.cuda.synchronize(); t3 = time.time()
    G = rng.standard_normal((B, N))
    st_saved = w.grad_fn.saved_tensors[1].cpu().numpy()
    w.backward(torch.tensor(G, dtype=torch.float32, device=dev))
    torch.cuda.synchronize(); t4 = time.time()
    Cn = tC.detach().double().cpu().numpy()
    w_ref, st, kkt, status = oracle.markowitz_fwd(tr.detach().double().cpu().numpy(), Cn, np.ones(B), np.ones(B), u)
    print("   saved-state mismatches vs oracle:", int((st_saved != st).sum()), "of", st.size)
    wk = w.detach().double().cpu().numpy()
    st_k = np.where(wk <= 0, -1, np.where(wk >= np.float32(u), 1, 0)).astype(np.int8)
    print("   state mismatches vs oracle:", int((st_k != st).sum()), "min |w_ref| among mismatches:", (np.abs(w_ref)[st_k != st].tolist() if (st_k != st).any() else None))
    d_r, d_C, d_g, d_a = oracle.markowitz_bwd(G, w_ref, st_k, Cn, np.ones(B), np.ones(B), u)
    rel = lambda a_, b_: np.abs(a_.double().cpu().numpy() - b_).max() / max(np.abs(b_).max(), 1e-30)
    print(f"B={B} L={L} N={N}: cov {t1-t0:.3f}s relerr {err_c:.2e} | fwd {t3-t2:.3f}s werr {np.abs(w.detach().double().cpu().numpy()-w_ref).max():.2e} status {layer.last_status.unique().tolist()} nfree {(st==0).sum(1).mean():.1f} | bwd {t4-t3:.3f}s d_rets {rel(tr.grad,d_r):.2e} d_C {rel(tC.grad,d_C):.2e} d_g {rel(g.grad,d_g):.2e} d_a {rel(a.grad,d_a):.2e}", flush=True)

"""GPU parity: CUDA kernels (through the layers, i.e. through the C ABI) vs the float64 CPU oracle
on identical seeded inputs, vs the committed golden fixtures, and property checks at full size.

Tolerances (north_star): weights 1e-5 abs, gradients 1e-4 rel (float32 kernels); float64 kernels are
held to 1e-9.  Sparsemax supports must match exactly.
"""
import os

import numpy as np
import pytest
import torch

import oracle
from deepdow_b200 import (CovarianceMatrix, NumericalMarkowitz, SoftmaxAllocator, SolverError,
                          SparsemaxAllocator)

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

W_ATOL = {torch.float32: 1e-5, torch.float64: 1e-9}
G_RTOL = {torch.float32: 1e-4, torch.float64: 1e-8}


def t(a, dtype, grad=False):
    return torch.tensor(np.asarray(a), dtype=dtype, device=DEV, requires_grad=grad)


def rel_err(a, b):
    a = a.detach().double().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
    b = np.asarray(b)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def make_problems(rng, B, N, kind="sym", alpha=1.0, rscale=0.5, gamma="ones"):
    rets = rng.standard_normal((B, N)) * rscale
    if kind == "sym":
        R = rng.standard_normal((B, N, N))
        C = R @ R.transpose(0, 2, 1) / N + 0.1 * np.eye(N)
    elif kind == "asym":
        C = rng.standard_normal((B, N, N)) * 0.5 + np.eye(N)
    elif kind == "dense":
        R = rng.standard_normal((B, N, N))
        C = R @ R.transpose(0, 2, 1) / N + 0.1 * np.eye(N)
        rets = rng.standard_normal((B, N)) * 0.02
    elif kind == "sqrtcov":
        X = rng.standard_normal((B, 40, N)) * 0.01
        C = oracle.covariance_fwd(X, 0.5, "diagonal", True)
        rets = rng.standard_normal((B, N)) * 0.01
    elif kind == "covsq":
        # BachelierNet's quirk (nn.py:142-144,171): the covariance ITSELF (sqrt=False) sits in the covmat_sqrt slot, so the
        # risk term is w' Sigma^2 w; with returns of the data's scale nearly every asset ends up free
        X = rng.standard_normal((B, 40, N)) * 0.01
        C = oracle.covariance_fwd(X, 0.5, "diagonal", False)
        rets = rng.standard_normal((B, N)) * 0.01
    g = np.ones(B) if gamma == "ones" else rng.uniform(0.5, 2.0, B)
    a = np.full(B, float(alpha)) * (rng.choice([-1.0, 1.0], B) if gamma != "ones" else 1.0)
    return rets, C, g, a


# ------------------------------------------------------------------ NumericalMarkowitz forward
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("B,N,u,kind,gamma", [
    (64, 20, 1.0, "sqrtcov", "ones"),      # config 1
    (37, 5, 1.0, "sym", "rand"),
    (33, 8, 0.3, "asym", "rand"),
    (64, 50, 1.0, "sym", "ones"),
    (128, 100, 0.2, "sym", "ones"),        # config 4 shape (reduced batch)
    (16, 100, 1.0, "asym", "rand"),
    (8, 129, 0.05, "sym", "ones"),         # 256-thread variant
    (4, 200, 0.1, "sym", "rand"),
    (4, 240, 1.0, "sqrtcov", "ones"),      # Q in global workspace, dense free set
    (3, 500, 0.01, "sym", "ones"),         # config 5 shape (reduced batch)
    # the diversified ("dense") regime the networks train in: most assets free, handled by the dense kernels
    (256, 100, 0.2, "dense", "ones"),      # config 4 shape, ~86 free (ThorpNet)
    (256, 50, 1.0, "covsq", "ones"),       # config 2: BachelierNet's Sigma^2 quirk
    (8, 129, 0.05, "dense", "ones"),       # 256-thread variant, dense free set
    (64, 100, 0.2, "dense", "rand"),       # dense + gamma tiling + signed alpha
])
def test_markowitz_forward_vs_oracle(dtype, B, N, u, kind, gamma):
    rng = np.random.default_rng(B * 1000 + N)
    rets, C, g, a = make_problems(rng, B, N, kind, 1.0, 0.5, gamma)
    w_ref, st_ref, kkt, status = oracle.markowitz_fwd(rets, C, g, a, u)
    assert kkt.max() < 1e-10 and (status == 0).all()
    layer = NumericalMarkowitz(N, max_weight=u)
    w = layer(t(rets, dtype), t(C, dtype), t(g, dtype), t(a, dtype))
    assert w.shape == (B, N) and w.dtype == dtype and w.device == torch.device(DEV)
    err = (w.double().cpu().numpy() - w_ref)
    assert np.abs(err).max() < W_ATOL[dtype], np.abs(err).max()
    assert torch.allclose(w.sum(1), torch.ones(B, dtype=dtype, device=DEV), atol=10 * W_ATOL[dtype])
    assert w.min() >= 0 and w.max() <= u + 1e-7
    assert int((layer.last_status & 12).ne(0).sum()) == 0
    assert_state_matches(w, w_ref, st_ref, u)
    if kind in ("dense", "covsq"):
        assert (st_ref == 0).sum(1).mean() > 0.5 * N      # the case is what it claims to be: most assets free


def assert_state_matches(w, w_ref, st_ref, u, tol=1e-6):
    """The active set the kernel reports (a function of the returned weights: 0 -> lower, u -> upper) equals the oracle's,
    except where the oracle's weight sits within `tol` of a bound (a degenerate asset)."""
    wn = w.detach().double().cpu().numpy()
    st = np.where(wn <= 0, -1, np.where(wn >= np.float64(w.new_tensor(u).item()), 1, 0))
    near = (np.abs(w_ref) < tol) | (np.abs(w_ref - u) < tol)
    assert ((st == st_ref) | near).all(), int(((st != st_ref) & ~near).sum())


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("fixture", ["markowitz_slsqp.npz", "markowitz_scipy_large.npz"])
def test_markowitz_golden_slsqp(dtype, fixture, golden_dir):
    z = np.load(os.path.join(golden_dir, fixture))
    for c in range(int(z["n_cases"])):
        k = f"m{c}"
        N = z[k + "_rets"].shape[1]
        layer = NumericalMarkowitz(N, max_weight=float(z[k + "_u"]))
        w = layer(t(z[k + "_rets"], dtype), t(z[k + "_C"], dtype), t(z[k + "_gamma"], dtype), t(z[k + "_alpha"], dtype))
        assert np.abs(w.double().cpu().numpy() - z[k + "_w"]).max() < 1e-5, c


def test_markowitz_reduces_to_sparsemax_goldens(golden_dir):
    z = np.load(os.path.join(golden_dir, "reference_vectors.npz"))
    x = z["sparsemax_known_x"]
    C = np.broadcast_to(np.eye(5), (2, 5, 5)).copy()
    w = NumericalMarkowitz(5)(t(2 * x, torch.float32), t(C, torch.float32), t(np.ones(2), torch.float32), t(np.zeros(2), torch.float32))
    assert np.allclose(w.cpu().numpy(), z["sparsemax_known_w"], atol=1e-3)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_markowitz_vertex_degenerate_and_edge_cases(dtype):
    rng = np.random.default_rng(11)
    # LP-like, exactly five assets at the cap (no free asset at the optimum)
    rets = rng.standard_normal((16, 30))
    C = np.broadcast_to(np.eye(30) * 1e-3, (16, 30, 30)).copy()
    g, a = np.ones(16), np.zeros(16)
    w_ref, *_ = oracle.markowitz_fwd(rets, C, g, a, 0.2)
    w = NumericalMarkowitz(30, 0.2)(t(rets, dtype), t(C, dtype), t(g, dtype), t(a, dtype))
    assert np.abs(w.double().cpu().numpy() - w_ref).max() < W_ATOL[dtype]
    # n_assets * max_weight == 1: a single feasible point
    w = NumericalMarkowitz(5, 0.2)(t(rets[:, :5], dtype), t(C[:, :5, :5], dtype), t(g, dtype), t(np.ones(16), dtype))
    assert torch.allclose(w, torch.full_like(w, 0.2))
    # infeasible box -> SolverError (the reference raises diffcp.SolverError)
    with pytest.raises(SolverError):
        NumericalMarkowitz(4, 0.2)(t(rets[:, :4], dtype), t(C[:, :4, :4], dtype), t(g, dtype), t(np.ones(16), dtype))
    # B = 1 with gamma of shape (1,) (examples/end_to_end/iid.py:362-365) and alpha = 0
    rets1, C1, _, _ = make_problems(rng, 1, 12)
    w_ref, *_ = oracle.markowitz_fwd(rets1, C1, np.ones(1) * 1.3, np.zeros(1), 1.0)
    w = NumericalMarkowitz(12)(t(rets1, dtype), t(C1, dtype), t([1.3], dtype), t([0.0], dtype))
    assert np.abs(w.double().cpu().numpy() - w_ref).max() < W_ATOL[dtype]


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_markowitz_backward_without_free_assets_and_with_many_capped(dtype):
    """Backward edge cases of the sparse-support kernel: (i) no free asset at the optimum (all weight on capped assets):
    every gradient is exactly zero; (ii) most of the weight on capped assets, a few free ones."""
    rng = np.random.default_rng(12)
    B, N, u = 16, 30, 0.2
    rets = rng.standard_normal((B, N))
    C = np.broadcast_to(np.eye(N) * 1e-3, (B, N, N)).copy()
    g, a = np.ones(B), np.zeros(B)
    G = rng.standard_normal((B, N))
    w_ref, st_ref, *_ = oracle.markowitz_fwd(rets, C, g, a, u)
    assert (st_ref == 0).sum() == 0                       # the case is what it claims to be
    ins = [t(rets, dtype, True), t(C, dtype, True), t(g, dtype, True), t(a, dtype, True)]
    w = NumericalMarkowitz(N, u)(*ins)
    grads = torch.autograd.grad(w, ins, t(G, dtype))
    assert all(float(x.abs().max()) == 0.0 for x in grads)
    # (ii) strong returns, tight cap: ~18 of 64 assets at the cap, a few free
    B, N, u = 24, 64, 0.05
    rets, C, g, a = make_problems(rng, B, N, "sym", 1.0, 3.0, "rand")
    G = rng.standard_normal((B, N))
    w_ref, st_ref, kkt, _ = oracle.markowitz_fwd(rets, C, g, a, u)
    assert (st_ref == 1).sum(1).min() >= 10
    ref = oracle.markowitz_bwd(G, w_ref, st_ref, C, g, a, u)
    ins = [t(rets, dtype, True), t(C, dtype, True), t(g, dtype, True), t(a, dtype, True)]
    layer = NumericalMarkowitz(N, u)
    w = layer(*ins)
    assert np.abs(w.detach().double().cpu().numpy() - w_ref).max() < W_ATOL[dtype]
    assert layer.solver_counters()[1] == 0                # stayed on the sparse-support kernel
    grads = torch.autograd.grad(w, ins, t(G, dtype))
    for name, got, want in zip(["rets", "covmat_sqrt", "gamma", "alpha"], grads, ref):
        assert rel_err(got, want) < G_RTOL[dtype], (name, rel_err(got, want))


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_markowitz_hard_problems_use_fallback_and_stay_exact(dtype):
    """Ill-conditioned Q where the active set can cycle: the interior-point fallback must kick in."""
    rng = np.random.default_rng(12)
    B, N = 256, 20
    V = np.linalg.qr(rng.standard_normal((B, N, N)))[0]
    lam = np.logspace(-2, 0.5, N)
    C = (V * lam) @ V.transpose(0, 2, 1)
    rets = rng.standard_normal((B, N))
    g, a = np.ones(B), np.zeros(B)
    for u in (0.2, 1.0):
        w_ref, st, kkt, status = oracle.markowitz_fwd(rets, C, g, a, u)
        layer = NumericalMarkowitz(N, u)
        w = layer(t(rets, dtype), t(C, dtype), t(g, dtype), t(a, dtype))
        tol = 2e-4 if dtype == torch.float32 else 1e-8   # cond(Q) ~ 1e5: float32 cannot give 1e-5 here
        assert np.abs(w.double().cpu().numpy() - w_ref).max() < tol
        assert int((layer.last_status & 12).ne(0).sum()) == 0


# ------------------------------------------------------------------ NumericalMarkowitz backward
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("B,N,u,kind,gamma", [
    (32, 7, 0.35, "asym", "rand"),
    (64, 20, 1.0, "sym", "ones"),
    (64, 50, 0.1, "sym", "rand"),
    (96, 100, 0.2, "sym", "ones"),
    (8, 150, 0.05, "sym", "ones"),
    (32, 100, 1.0, "dense", "ones"),       # nearly all assets free: generic factorisation + streaming path
    (64, 100, 0.2, "dense", "rand"),       # config 4's regime with the gamma tiling
    (64, 50, 1.0, "covsq", "ones"),        # config 2's regime (BachelierNet)
    (4, 240, 0.05, "dense", "ones"),       # factor in global workspace
    (3, 500, 0.01, "sym", "rand"),         # config 5 shape (reduced batch)
])
def test_markowitz_backward_vs_oracle(dtype, B, N, u, kind, gamma):
    rng = np.random.default_rng(B * 7 + N)
    rets, C, g, a = make_problems(rng, B, N, kind, 1.0, 0.5, gamma)
    G = rng.standard_normal((B, N))
    w_ref, st_ref, kkt, _ = oracle.markowitz_fwd(rets, C, g, a, u)
    ref = oracle.markowitz_bwd(G, w_ref, st_ref, C, g, a, u)
    tr, tC, tg, ta = t(rets, dtype, True), t(C, dtype, True), t(g, dtype, True), t(a, dtype, True)
    w = NumericalMarkowitz(N, u)(tr, tC, tg, ta)
    assert np.abs(w.detach().double().cpu().numpy() - w_ref).max() < W_ATOL[dtype]
    assert_state_matches(w, w_ref, st_ref, u)
    (w * t(G, dtype)).sum().backward()
    for name, got, want in zip(["rets", "covmat_sqrt", "gamma", "alpha"], [tr.grad, tC.grad, tg.grad, ta.grad], ref):
        assert got is not None and got.shape == want.shape, name
        assert rel_err(got, want) < G_RTOL[dtype], (name, rel_err(got, want))


def test_markowitz_status_modes():
    rng = np.random.default_rng(4)
    rets, C, g, a = make_problems(rng, 16, 12)
    args = [t(x, torch.float32) for x in (rets, C, g, a)]
    ref = None
    for mode in ("lazy", True, False):
        layer = NumericalMarkowitz(12, 0.5)
        layer.check_status = mode
        for _ in range(12):                     # more calls than status slots: the ring must recycle
            w = layer(*args)
        layer.synchronize_status()
        assert int(layer.last_status.ne(0).sum()) == 0
        ref = w if ref is None else ref
        assert torch.equal(w, ref)
    # an infeasible box is rejected on the host, immediately, in every checking mode
    for mode in ("lazy", True):
        layer = NumericalMarkowitz(12, 0.05)
        layer.check_status = mode
        with pytest.raises(SolverError):
            layer(*args)


def test_markowitz_no_grad_and_double_backward_safety():
    rng = np.random.default_rng(3)
    rets, C, g, a = make_problems(rng, 8, 10)
    with torch.no_grad():
        w = NumericalMarkowitz(10)(t(rets, torch.float32), t(C, torch.float32), t(g, torch.float32), t(a, torch.float32))
    assert not w.requires_grad


def test_markowitz_full_size_properties():
    """Config 4 at full size (B=4096, N=100, max_weight=0.2): KKT certificate computed in float64 on the GPU."""
    B, N, u = 4096, 100, 0.2
    gen = torch.Generator(device=DEV).manual_seed(4000)
    R = torch.randn(B, N, N, generator=gen, device=DEV)
    C = R @ R.transpose(1, 2) / N + 0.1 * torch.eye(N, device=DEV)
    rets = torch.randn(B, N, generator=gen, device=DEV) * 0.5
    g = torch.ones(B, device=DEV)
    a = torch.ones(B, device=DEV)
    layer = NumericalMarkowitz(N, u)
    w = layer(rets, C, g, a)
    assert int(layer.last_status.ne(0).sum()) == 0
    wd, Cd = w.double(), C.double()
    Q = Cd.transpose(1, 2) @ Cd + torch.eye(N, device=DEV, dtype=torch.float64)
    h = 2 * (Q @ wd[..., None])[..., 0] - rets.double()
    free = (wd > 1e-7) & (wd < u - 1e-7)
    nu = -(h * free).sum(1) / free.sum(1).clamp(min=1)
    grad = h + nu[:, None]
    res = torch.where(free, grad.abs(), torch.where(wd <= 1e-7, (-grad).clamp(min=0), grad.clamp(min=0)))
    assert (wd.sum(1) - 1).abs().max() < 1e-5
    assert wd.min() >= 0 and wd.max() <= u + 1e-7
    assert res.max() < 2e-4, res.max()      # gradient units; |h| ~ 1..10
    # sample of problems against the oracle
    idx = torch.arange(0, B, 64)
    w_ref, *_ = oracle.markowitz_fwd(rets[idx].cpu().numpy(), C[idx].cpu().numpy(), np.ones(len(idx)), np.ones(len(idx)), u)
    assert np.abs(w[idx].double().cpu().numpy() - w_ref).max() < 1e-5


def test_markowitz_config5_full_size_properties():
    """Config 5 per-GPU share at full size (B=1024, N=500, lookback 250): CovarianceMatrix(sqrt) -> NumericalMarkowitz
    through the dense kernels with Q in the global workspace; KKT certificate in float64 on the GPU, fwd + bwd."""
    B, L, N, u = 1024, 250, 500, 0.05
    gen = torch.Generator(device=DEV).manual_seed(5000)
    x = torch.randn(B, L, N, generator=gen, device=DEV) * 0.01
    C = CovarianceMatrix(sqrt=True, shrinkage_strategy="diagonal")(x)
    rets = torch.randn(B, N, generator=gen, device=DEV) * 0.01
    g = torch.ones(B, device=DEV)
    a = torch.ones(B, device=DEV)
    layer = NumericalMarkowitz(N, u)
    ins = [t.requires_grad_() for t in (rets, C.detach().clone(), g, a)]
    w = layer(*ins)
    layer.synchronize_status()
    assert int((layer.last_status & 12).ne(0).sum()) == 0        # nothing infeasible / inexact
    wd, Cd = w.detach().double(), ins[1].detach().double()
    Aw = (Cd @ wd[..., None])[..., 0]
    h = 2 * ((Cd.transpose(1, 2) @ Aw[..., None])[..., 0] + wd) - ins[0].detach().double()
    free = (wd > 1e-7) & (wd < u - 1e-7)
    nu = -(h * free).sum(1) / free.sum(1).clamp(min=1)
    grad = h + nu[:, None]
    res = torch.where(free, grad.abs(), torch.where(wd <= 1e-7, (-grad).clamp(min=0), grad.clamp(min=0)))
    assert (wd.sum(1) - 1).abs().max() < 1e-5
    assert wd.min() >= 0 and wd.max() <= u + 1e-7
    assert res.max() < 1e-4, res.max()
    G = torch.randn(B, N, generator=gen, device=DEV)
    grads = torch.autograd.grad(w, ins, G)
    assert all(torch.isfinite(t).all() for t in grads)
    # the adjoint of a budget-preserving map: d_rets sums to zero over each portfolio (1'd = 0)
    assert grads[0].double().sum(1).abs().max() < 1e-4 * grads[0].abs().max().clamp(min=1.0)
    # sample of problems against the oracle
    idx = torch.arange(0, B, 256)
    w_ref, *_ = oracle.markowitz_fwd(ins[0][idx].detach().cpu().numpy(), ins[1][idx].detach().cpu().numpy(),
                                     np.ones(len(idx)), np.ones(len(idx)), u)
    assert np.abs(w[idx].detach().double().cpu().numpy() - w_ref).max() < 1e-5


# ------------------------------------------------------------------ sparsemax / softmax
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("B,N,u,scale", [(1024, 100, 1.0, 1.0), (1024, 100, 0.05, 10.0), (33, 5, 0.25, 1.0),
                                         (7, 300, 0.013, 3.0), (5, 1000, 1.0, 1.0),
                                         # one case per (lanes per row, vector / scalar) layout of rowproj.cu, ragged batch ends
                                         (37, 50, 0.1, 5.0), (1, 20, 0.2, 10.0), (130, 64, 0.03, 20.0), (19, 200, 0.05, 8.0),
                                         (9, 130, 0.3, 4.0), (3, 2048, 0.001, 30.0), (11, 513, 1.0, 2.0)])
def test_sparsemax_vs_oracle(dtype, B, N, u, scale):
    rng = np.random.default_rng(N)
    u = float(np.float32(u)) if dtype == torch.float32 else u   # the cap exactly as the kernel sees it
    x = rng.uniform(-0.5, 0.5, (B, N)) * scale
    temp = rng.uniform(0.5, 2.0, B)
    G = rng.standard_normal((B, N))
    xin = t(x, dtype).double().cpu().numpy()       # the exact inputs the kernel sees
    tin = t(temp, dtype).double().cpu().numpy()
    w_ref = oracle.capped_simplex_fwd(xin, tin, u)
    tx, tt = t(x, dtype, True), t(temp, dtype, True)
    w = SparsemaxAllocator(N, temperature=None, max_weight=u)(tx, tt)
    wn = w.detach().double().cpu().numpy()
    assert np.abs(wn - w_ref).max() < W_ATOL[dtype]
    # supports must match exactly, except entries the oracle itself puts within rounding of a bound
    margin = 1e-6 if dtype == torch.float32 else 1e-12
    sup_ref, sup = w_ref > 0, wn > 0
    mism = sup_ref != sup
    # a mismatch is tolerated only where the oracle's own weight is within rounding of zero
    assert not (mism & (np.abs(w_ref) > margin)).any()
    near_zero_boundary = mism.sum()
    assert near_zero_boundary <= max(1, B // 200)
    (w * t(G, dtype)).sum().backward()
    dx_ref, dt_ref = oracle.capped_simplex_bwd(G, wn, xin, tin, u)
    assert rel_err(tx.grad, dx_ref) < G_RTOL[dtype]
    assert rel_err(tt.grad, dt_ref) < 10 * G_RTOL[dtype]


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("B,N,u,scale", [(1024, 100, 1.0, 1.0), (1024, 100, 0.05, 10.0), (33, 5, 0.25, 1.0), (6, 700, 0.01, 3.0),
                                         (37, 50, 0.1, 5.0), (1, 20, 0.2, 10.0), (130, 64, 0.03, 20.0), (19, 200, 0.05, 8.0),
                                         (9, 130, 0.3, 4.0), (3, 2048, 0.001, 30.0), (11, 513, 1.0, 2.0)])
def test_softmax_vs_oracle(dtype, B, N, u, scale):
    rng = np.random.default_rng(N + 1)
    u = float(np.float32(u)) if dtype == torch.float32 else u
    x = rng.uniform(-0.5, 0.5, (B, N)) * scale
    temp = rng.uniform(0.5, 2.0, B)
    G = rng.standard_normal((B, N))
    xin = t(x, dtype).double().cpu().numpy()
    tin = t(temp, dtype).double().cpu().numpy()
    w_ref = oracle.capped_softmax_fwd(xin, tin, u)
    tx, tt = t(x, dtype, True), t(temp, dtype, True)
    w = SoftmaxAllocator(temperature=None, formulation="variational", n_assets=N, max_weight=u)(tx, tt)
    wn = w.detach().double().cpu().numpy()
    assert np.abs(wn - w_ref).max() < W_ATOL[dtype]
    (w * t(G, dtype)).sum().backward()
    dx_ref, dt_ref = oracle.capped_softmax_bwd(G, wn, xin, tin, u)
    assert rel_err(tx.grad, dx_ref) < G_RTOL[dtype]
    assert rel_err(tt.grad, dt_ref) < 10 * G_RTOL[dtype]


def test_row_allocators_reference_goldens(golden_dir):
    z = np.load(os.path.join(golden_dir, "reference_vectors.npz"))
    f32 = torch.float32
    # tests/test_layers.py:821-838
    w = SparsemaxAllocator(5, temperature=1)(t(z["sparsemax_known_x"], f32))
    assert torch.allclose(w, t(z["sparsemax_known_w"], f32), atol=1e-3)
    # docs/source/layers.rst:381-394
    w = SparsemaxAllocator(3, temperature=1)(t(z["sparsemax_doc_x"], f32))
    assert torch.allclose(w, t(z["sparsemax_doc_w"], f32), atol=2e-4)
    # tests/test_layers.py:815-819, 752-759
    assert torch.allclose(SparsemaxAllocator(5, temperature=1)(torch.ones(2, 5, device=DEV)), torch.full((2, 5), 0.2, device=DEV))
    assert torch.allclose(SoftmaxAllocator(formulation="variational", n_assets=5, temperature=1)(torch.ones(2, 5, device=DEV)),
                          torch.full((2, 5), 0.2, device=DEV), atol=1e-4)
    # tests/test_layers.py:761-776, 840-850
    for u in z["capped_u"]:
        x = t(z["capped_x"], f32)
        for layer_c, layer_u in [(SparsemaxAllocator(5, temperature=1, max_weight=float(u)), SparsemaxAllocator(5, temperature=1)),
                                 (SoftmaxAllocator(n_assets=5, temperature=1, max_weight=float(u), formulation="variational"),
                                  SoftmaxAllocator(n_assets=5, temperature=1, formulation="variational"))]:
            wc, wu = layer_c(x), layer_u(x)
            assert not torch.allclose(wc, wu)
            assert wc.max().item() == pytest.approx(float(u), abs=1e-6)
    # tests/test_layers.py:726-750: variational == analytical softmax
    torch.manual_seed(2)
    rets = (torch.randint(10, size=(3, 6)) + torch.rand(size=(3, 6))).to(DEV)
    wa = SoftmaxAllocator(formulation="analytical")(rets)
    wv = SoftmaxAllocator(formulation="variational", n_assets=6)(rets)
    assert wa.shape == wv.shape and wa.dtype == wv.dtype and wa.device == wv.device
    assert torch.allclose(wa, wv, atol=1e-5)


# ------------------------------------------------------------------ CovarianceMatrix
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_covariance_vs_reference_fixtures(dtype, golden_dir):
    z = np.load(os.path.join(golden_dir, "covariance_reference.npz"))
    for c in range(int(z["n_cases"])):
        k = f"c{c}"
        strategy, sqrt = z[k + "_meta"]
        strategy = None if strategy == "None" else str(strategy)
        sqrt = bool(int(sqrt))
        x, coef = t(z[k + "_x"], dtype, True), t(z[k + "_coef"], dtype, True)
        layer = CovarianceMatrix(sqrt=sqrt, shrinkage_strategy=strategy, shrinkage_coef=None)
        y = layer(x, coef)
        ftol = 2e-5 if dtype == torch.float32 else 1e-10
        assert rel_err(y, z[k + "_y"]) < ftol, (c, strategy, sqrt, rel_err(y, z[k + "_y"]))
        (y * t(z[k + "_G"], dtype)).sum().backward()
        if not np.isfinite(z[k + "_dx"]).all():
            dx_ref, dc_ref = oracle.covariance_bwd(z[k + "_G"], z[k + "_x"], z[k + "_coef"], strategy, sqrt)
        else:
            dx_ref, dc_ref = z[k + "_dx"], z[k + "_dcoef"]
        gtol = 2e-3 if dtype == torch.float32 else 1e-7     # sqrt backward divides by sqrt-eigenvalue sums
        assert rel_err(x.grad, dx_ref) < gtol, (c, strategy, sqrt, rel_err(x.grad, dx_ref))
        if strategy is not None:
            assert rel_err(coef.grad, dc_ref) < gtol, (c, strategy, sqrt, rel_err(coef.grad, dc_ref))


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("B,L,N", [(64, 40, 20), (256, 40, 50), (16, 40, 100), (4, 60, 130), (2, 250, 200),
                                   # beyond shared memory: tensor-core Gram + blocked Jacobi in float32 (eig_block.cu), padded blocks
                                   (3, 300, 260), (2, 250, 500), (2, 40, 385),
                                   # odd shapes: no 16-byte alignment for the staged tile / the flat store of the plain kernel
                                   (5, 7, 3), (3, 2, 9), (9, 33, 51), (300, 3, 1)])
def test_covariance_vs_oracle_sizes(dtype, B, L, N):
    rng = np.random.default_rng(L + N)
    x = rng.standard_normal((B, L, N)) * 0.01
    xin = t(x, dtype).double().cpu().numpy()
    for strategy in ["diagonal", "scaled_identity", "identity", None]:
        for sqrt in [False, True]:
            if sqrt and strategy in ("identity", None):
                continue        # the square root of these is covered by the reference fixtures
            y = CovarianceMatrix(sqrt=sqrt, shrinkage_strategy=strategy)(t(x, dtype))
            ref = oracle.covariance_fwd(xin, 0.5, strategy, sqrt)
            tol = 3e-5 if dtype == torch.float32 else 1e-10
            assert rel_err(y, ref) < tol, (strategy, sqrt, rel_err(y, ref))


def test_covariance_reference_behaviour(golden_dir):
    z = np.load(os.path.join(golden_dir, "covariance_reference.npz"))
    # tests/test_layers.py:203-209
    with pytest.raises(ValueError):
        CovarianceMatrix(shrinkage_strategy="fake")
    with pytest.raises(ValueError):
        CovarianceMatrix(shrinkage_coef=None)(torch.ones(1, 5, 2, device=DEV))
    # :221-225 dim == 1
    with pytest.raises(ZeroDivisionError):
        CovarianceMatrix()(torch.ones(2, 1, 3, device=DEV))
    # :243-256 and docs/source/layers.rst:474-487
    xd = t(z["doc_x"], torch.float32)
    cov = CovarianceMatrix(sqrt=False)(xd)
    sq = CovarianceMatrix(sqrt=True)(xd)
    assert torch.allclose(cov, t(z["doc_cov"], torch.float32), rtol=1e-4)
    assert torch.allclose(sq, t(z["doc_sqrt"], torch.float32), rtol=1e-3, atol=1e-3)
    assert torch.allclose(cov[0], sq[0] @ sq[0], atol=1e-2)
    # rank deficient input (strategy None, dim - 1 < n_assets)
    y = CovarianceMatrix(sqrt=True, shrinkage_strategy=None, shrinkage_coef=1.0)(t(z["rd_x"], torch.float64))
    assert rel_err(y, z["rd_y"]) < 1e-7
    # :258-289 constructor coefficient == per-sample tensor coefficient
    x = torch.rand(3, 4, 5, device=DEV) * 100
    l1 = CovarianceMatrix(sqrt=False, shrinkage_strategy="diagonal", shrinkage_coef=0.3)
    l2 = CovarianceMatrix(sqrt=False, shrinkage_strategy="diagonal", shrinkage_coef=None)
    assert torch.allclose(l1(x), l2(x, 0.3 * torch.ones(3, device=DEV)))
    assert sum(p.numel() for p in l1.parameters()) == 0


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_compute_sqrt_static(dtype):
    """CovarianceMatrix.compute_sqrt (misc.py:168-201) on a given matrix, forward and gradient."""
    rng = np.random.default_rng(21)
    for N in (5, 40, 110, 300):
        R = rng.standard_normal((N, N))
        M = R @ R.T / N + 0.05 * np.eye(N)
        G = rng.standard_normal((N, N))
        m = t(M, dtype, True)
        r = CovarianceMatrix.compute_sqrt(m)
        Min = m.detach().double().cpu().numpy()
        assert rel_err(r, oracle.psd_sqrt_fwd(Min[None])[0]) < (2e-5 if dtype == torch.float32 else 1e-10)
        assert rel_err(r @ r, Min) < (5e-5 if dtype == torch.float32 else 1e-10)
        (r * t(G, dtype)).sum().backward()
        assert rel_err(m.grad, oracle.psd_sqrt_bwd(G[None], Min[None])[0]) < (2e-3 if dtype == torch.float32 else 1e-7)


# ------------------------------------------------------------------ row allocators: randomized sweep over every layout
@pytest.mark.parametrize("seed", range(6))
def test_row_allocators_random_shapes(seed):
    """Random (n_samples, n_assets, cap, scale) per seed: every lanes-per-row / vector / scalar variant of rowproj.cu, ragged batch
    ends, caps from barely feasible to loose, logits from nearly flat to widely spread; sparsemax, capped softmax and the greedy LP."""
    from deepdow_b200.benchmarks import MaximumReturn
    rng = np.random.default_rng(1000 + seed)
    for _ in range(12):
        N = int(rng.choice([2, 3, 4, 7, 16, 31, 64, 65, 100, 128, 129, 200, 256, 257, 500, 512, 600, 1024, 1500]))
        B = int(rng.integers(1, 70))
        slack = float(rng.choice([1.0 + 1e-3, 1.3, 2.0, 7.7, N]))          # n_assets * cap
        u = min(1.0, slack / N)
        if N * float(np.float32(u)) < 1:
            u = min(1.0, 1.01 * slack / N)
        scale = float(rng.choice([1e-3, 0.3, 3.0, 40.0]))
        dtype = torch.float32 if rng.random() < 0.6 else torch.float64
        uk = float(np.float32(u)) if dtype == torch.float32 else u
        x = rng.standard_normal((B, N)) * scale
        xin = t(x, dtype).double().cpu().numpy()
        tag = (N, B, u, scale, str(dtype))
        w = SparsemaxAllocator(N, temperature=1, max_weight=uk)(t(x, dtype)).double().cpu().numpy()
        ref = oracle.capped_simplex_fwd(xin, np.ones(B), uk)
        assert np.abs(w - ref).max() < W_ATOL[dtype] * max(1.0, scale / 3), ("sparsemax",) + tag
        assert np.abs(w.sum(1) - 1).max() < 1e-5 and w.min() >= 0 and w.max() <= uk * (1 + 1e-6), ("sparsemax",) + tag
        w = SoftmaxAllocator(temperature=1, formulation="variational", n_assets=N, max_weight=uk)(t(x, dtype)).double().cpu().numpy()
        ref = oracle.capped_softmax_fwd(xin, np.ones(B), uk)
        assert np.abs(w - ref).max() < W_ATOL[dtype] * max(1.0, scale / 3), ("softmax",) + tag
        w = MaximumReturn(max_weight=uk)(t(x, dtype)[:, None, None, :]).double().cpu().numpy()
        ref = oracle.capped_argmax_fwd(xin, uk)
        assert np.abs(w - ref).max() < (2e-6 if dtype == torch.float32 else 1e-12), ("argmax",) + tag


# ------------------------------------------------------------------ NumericalMarkowitz: randomized sweep
@pytest.mark.parametrize("seed", range(5))
def test_markowitz_random_shapes(seed):
    """Random (batch, n_assets, cap, problem family, gamma, alpha, return scale) per seed, float64 kernels against the oracle:
    forward at 1e-8, every gradient at 1e-6 relative.  Covers the kernel hand-overs (sparse-support -> dense -> interior point),
    ragged sizes on both sides of every dispatch boundary, caps from barely feasible to absent."""
    rng = np.random.default_rng(500 + seed)
    for _ in range(7):
        N = int(rng.choice([2, 3, 5, 8, 9, 17, 33, 48, 49, 64, 100, 127, 128, 129, 160, 231, 300]))
        B = int(rng.integers(1, 24 if N < 200 else 5))
        slack = float(rng.choice([1.05, 1.5, 3.0, N / 2 + 1, N]))
        u = min(1.0, slack / N)
        kind = str(rng.choice(["sym", "asym", "dense"]))
        alpha = float(rng.choice([1e-2, 1.0, 10.0])) if kind == "asym" else float(rng.choice([0.0, 1e-2, 1.0]))
        rscale = float(rng.choice([0.02, 0.5, 5.0]))
        gamma = str(rng.choice(["ones", "rand"]))
        rets, C, g, a = make_problems(rng, B, N, kind, alpha, rscale, gamma)
        if kind == "dense":
            rets = rng.standard_normal((B, N)) * rscale * 0.04
        tag = (N, B, u, kind, alpha, rscale, gamma)
        w_ref, st_ref, kkt, status = oracle.markowitz_fwd(rets, C, g, a, u)
        assert kkt.max() < 1e-9 and (status == 0).all(), tag
        G = rng.standard_normal((B, N))
        ref = oracle.markowitz_bwd(G, w_ref, st_ref, C, g, a, u)
        dt = torch.float64
        tr, tC, tg, ta = t(rets, dt, True), t(C, dt, True), t(g, dt, True), t(a, dt, True)
        w = NumericalMarkowitz(N, u)(tr, tC, tg, ta)
        assert np.abs(w.detach().cpu().numpy() - w_ref).max() < 1e-8, tag
        # weights exactly on a bound with a vanishing multiplier make the gradient non-unique: skip the gradient check there
        wn = w_ref
        near = (np.minimum(wn, u - wn) < 1e-9) & (st_ref == 0)
        if near.any():
            continue
        (w * t(G, dt)).sum().backward()
        for name, got, want in zip(["rets", "covmat_sqrt", "gamma", "alpha"], [tr.grad, tC.grad, tg.grad, ta.grad], ref):
            assert rel_err(got, want) < 1e-6, (name, rel_err(got, want)) + tag
        # float32 kernels on the same problems: the north-star tolerances, for the well-conditioned family
        if kind == "sym" and alpha >= 1e-2:
            w32 = NumericalMarkowitz(N, float(np.float32(u)) if N * float(np.float32(u)) >= 1 else u)(
                t(rets, torch.float32), t(C, torch.float32), t(g, torch.float32), t(a, torch.float32))
            assert np.abs(w32.double().cpu().numpy() - w_ref).max() < 3e-5 * max(1.0, rscale), tag


# ------------------------------------------------------------------ CovarianceMatrix: randomized sweep
@pytest.mark.parametrize("seed", range(4))
def test_covariance_random_shapes(seed):
    """Random (batch, lookback, n_assets, strategy, sqrt, per-sample coefficient, data scale) per seed against the oracle: float64
    forward and gradients (x and shrinkage_coef), float32 forward.  Crosses the plain-kernel / general-kernel boundary (75 KB of
    shared memory), aligned and unaligned tile sizes, lookbacks shorter and longer than the asset count."""
    rng = np.random.default_rng(900 + seed)
    for _ in range(8):
        N = int(rng.choice([1, 2, 3, 5, 8, 13, 20, 32, 50, 51, 64, 100, 101, 130]))
        L = int(rng.choice([2, 3, 7, 16, 40, 41, 60, 120]))
        B = int(rng.integers(1, 12 if N < 100 else 4))
        strategy = [None, "diagonal", "identity", "scaled_identity"][int(rng.integers(0, 4))]
        sqrt = bool(rng.integers(0, 2))
        scale = float(rng.choice([0.01, 1.0, 30.0]))
        x = rng.standard_normal((B, L, N)) * scale + rng.standard_normal((1, 1, N)) * scale   # non-zero column means
        coef = rng.uniform(0.2, 0.9, B)
        G = rng.standard_normal((B, N, N))
        tag = (N, L, B, strategy, sqrt, scale)
        ref = oracle.covariance_fwd(x, coef, strategy, sqrt)
        layer = CovarianceMatrix(sqrt=sqrt, shrinkage_strategy=strategy, shrinkage_coef=None if strategy is not None else 0.5)
        tx = t(x, torch.float64, True)
        tc = t(coef, torch.float64, True) if strategy is not None else None
        y = layer(tx, tc) if strategy is not None else layer(tx)
        assert rel_err(y, ref) < 1e-9, tag
        # gradients: only where the square root is smooth (full-rank shrunk covariance) or no square root is taken
        full_rank = strategy in ("diagonal", "identity", "scaled_identity") or L - 1 >= N
        if not sqrt or (full_rank and strategy is not None):
            (y * t(G, torch.float64)).sum().backward()
            dx_ref, dc_ref = oracle.covariance_bwd(G, x, coef, strategy, sqrt)
            assert rel_err(tx.grad, dx_ref) < 1e-6, ("dx", rel_err(tx.grad, dx_ref)) + tag
            if strategy is not None:
                assert rel_err(tc.grad, dc_ref) < 1e-6, ("dcoef", rel_err(tc.grad, dc_ref)) + tag
        x32 = t(x, torch.float32)
        if sqrt:
            # the reference drops singular values below s.max * n * eps OF THE INPUT DTYPE (misc.py:185-187): a spectrum that
            # reaches down to the float32 threshold is cut differently in float32 and in the float64 oracle -- not comparable
            lam = np.abs(np.linalg.eigvalsh(oracle.covariance_fwd(x, coef, strategy, False)))
            if (lam.min(1) < 1e3 * lam.max(1) * N * np.finfo(np.float32).eps).any():
                continue
        ref32 = oracle.covariance_fwd(x32.double().cpu().numpy(), coef.astype(np.float32).astype(np.float64), strategy, sqrt)
        y32 = layer(x32, t(coef, torch.float32)) if strategy is not None else layer(x32)
        assert rel_err(y32, ref32) < (1e-4 if sqrt else 3e-5), ("f32", rel_err(y32, ref32)) + tag


# ------------------------------------------------------------------ repeatability (a proxy for shared-memory races)
def test_kernels_are_bitwise_repeatable():
    """compute-sanitizer's racecheck is not available on the GPU pool; a missing barrier in the blocked LDL', the grouped row
    kernels or the staged covariance kernel shows up as run-to-run differences, so every output is required to be bit-identical
    over repeated launches on the same inputs (dense-regime Markowitz at three sizes around the panel width, rows, covariance)."""
    gen = torch.Generator(device=DEV).manual_seed(11)
    for N, B, u in ((100, 64, 0.2), (61, 33, 1.0), (130, 9, 0.05), (250, 3, 0.02)):
        R = torch.randn(B, N, N, generator=gen, device=DEV)
        C = (R @ R.transpose(1, 2) / N + 0.1 * torch.eye(N, device=DEV)).contiguous().requires_grad_()
        rets = (torch.randn(B, N, generator=gen, device=DEV) * 0.02).requires_grad_()
        o = torch.ones(B, device=DEV)
        G = torch.randn(B, N, generator=gen, device=DEV)
        layer = NumericalMarkowitz(N, u)
        ref = None
        for _ in range(6):
            w = layer(rets, C, o, o)
            gr, gC = torch.autograd.grad(w, (rets, C), G)
            cur = (w.detach().clone(), gr.clone(), gC.clone())
            if ref is None:
                ref = cur
                assert ((w > 0) & (w < u)).sum(1).float().mean() > 0.5 * min(N, 1 / u)      # the dense path is the one exercised
            else:
                assert all(torch.equal(a, b) for a, b in zip(ref, cur)), (N, B, u)
    x = torch.randn(777, 100, generator=gen, device=DEV) * 3
    xc = torch.randn(99, 40, 50, generator=gen, device=DEV)
    outs = None
    for _ in range(6):
        cur = (SparsemaxAllocator(100, 1, 0.05)(x), SoftmaxAllocator(1, "variational", 100, 0.05)(x), CovarianceMatrix(sqrt=False)(xc),
               CovarianceMatrix(sqrt=True)(xc))
        if outs is None:
            outs = cur
        else:
            assert all(torch.equal(a, b) for a, b in zip(outs, cur))


# ------------------------------------------------------------------ Resample over the CUDA layers (allocate.py:276-411)
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("sqrt", [False, True])
def test_resample_over_numerical_markowitz(dtype, sqrt):
    """tests/test_layers.py:532-594 (NumericalMarkowitz arm): shape / dtype / device, the same seed gives the same weights,
    gradients reach ``rets`` -- and here also: the weights and ``rets.grad`` equal the oracle's on the very draws the layer used
    (the chain rsample -> CovarianceMatrix(sqrt=True) -> NumericalMarkowitz; the centred covariance does not depend on the
    location, so d w / d rets goes through the mean of the draws only)."""
    from deepdow_b200 import Resample
    n_samples, n_assets, n_draws, n_portfolios, u, seed = 3, 6, 12, 4, 0.6, 7
    rng = np.random.default_rng(31)
    R = rng.standard_normal((n_samples, n_assets, n_assets)) * 0.05
    cov = R @ R.transpose(0, 2, 1) + 0.01 * np.eye(n_assets)
    matrix = oracle.psd_sqrt_fwd(cov) if sqrt else cov
    rets = t(rng.standard_normal((n_samples, n_assets)) * 0.05, dtype, True)
    mat = t(matrix, dtype, True)
    gamma, alpha = t(np.ones(n_samples), dtype), t(np.ones(n_samples), dtype)
    with pytest.raises(TypeError):
        Resample(SparsemaxAllocator(n_assets))
    layer = Resample(NumericalMarkowitz(n_assets, max_weight=u), n_draws=n_draws, n_portfolios=n_portfolios, sqrt=sqrt, random_state=seed)
    w = layer(mat, rets, gamma=gamma, alpha=alpha)
    assert w.shape == (n_samples, n_assets) and w.dtype == dtype and w.device == torch.device(DEV)
    assert torch.allclose(w.sum(1), torch.ones(n_samples, dtype=dtype, device=DEV), atol=1e-5)
    w2 = layer(mat, rets, gamma=gamma, alpha=alpha)
    assert torch.equal(w, w2)                                  # random_state: same draws, same weights
    G = rng.standard_normal((n_samples, n_assets))
    (w * t(G, dtype)).sum().backward()
    assert rets.grad is not None and torch.isfinite(rets.grad).all()
    assert mat.grad is not None and torch.isfinite(mat.grad).all() and float(mat.grad.abs().max()) > 0
    # replay the draws (same seed, same device, same call order) and push them through the oracle
    torch.manual_seed(seed)
    with torch.no_grad():
        covmat = mat @ mat if sqrt else mat
        dist = torch.distributions.MultivariateNormal(loc=rets.detach(), covariance_matrix=covmat)
        w_ref = np.zeros((n_samples, n_assets)); d_ref = np.zeros((n_samples, n_assets))
        for _ in range(n_portfolios):
            draws = dist.rsample((n_draws,)).permute(1, 0, 2).double().cpu().numpy()       # (n_samples, n_draws, n_assets)
            C = oracle.covariance_fwd(draws, 0.5, "diagonal", True)
            wp, stp, kkt, _ = oracle.markowitz_fwd(draws.mean(1), C, np.ones(n_samples), np.ones(n_samples), u)
            w_ref += wp / n_portfolios
            d_ref += oracle.markowitz_bwd(G / n_portfolios, wp, stp, C, np.ones(n_samples), np.ones(n_samples), u)[0]
    tolw = 2e-5 if dtype == torch.float32 else 1e-9
    assert np.abs(w.detach().double().cpu().numpy() - w_ref).max() < tolw
    assert rel_err(rets.grad, d_ref) < (2e-3 if dtype == torch.float32 else 1e-7)


# ------------------------------------------------------------------ NumericalRiskBudgeting (allocate.py:631-691)
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("B,N,u,kind", [(16, 6, 1.0, "ref"), (64, 20, 1.0, "ref"), (32, 50, 0.05, "ref"), (64, 100, 1.0, "sqrtcov"),
                                        (32, 100, 0.02, "sqrtcov"), (4, 150, 0.02, "ref"), (3, 300, 1.0, "sqrtcov")])
def test_risk_budgeting_vs_oracle(dtype, B, N, u, kind):
    from deepdow_b200 import NumericalRiskBudgeting
    rng = np.random.default_rng(B * 13 + N)
    if kind == "ref":                        # the matrices of tests/test_layers.py:1011-1015
        R = rng.random((B, N, N))
        A = R @ R / N + np.eye(N)
    else:
        A = oracle.covariance_fwd(rng.standard_normal((B, max(40, N // 2), N)), 0.5, "diagonal", True)
    b = rng.random((B, N)); b /= b.sum(1, keepdims=True)
    G = rng.standard_normal((B, N))
    tA, tb = t(A, dtype, True), t(b, dtype, True)
    w_ref, st_ref, kkt, status = oracle.risk_budgeting_fwd(tA.detach().double().cpu().numpy(), tb.detach().double().cpu().numpy(), u)
    assert kkt.max() < 1e-8 and (status == 0).all()
    layer = NumericalRiskBudgeting(N, max_weight=u)
    w = layer(tA, tb)
    assert w.shape == (B, N) and w.dtype == dtype and w.device == torch.device(DEV)
    assert np.abs(w.detach().double().cpu().numpy() - w_ref).max() < W_ATOL[dtype]
    assert torch.allclose(w.sum(1), torch.ones(B, dtype=dtype, device=DEV), atol=10 * W_ATOL[dtype])
    assert w.min() > 0 and w.max() <= u + 1e-7
    assert int((layer.last_status & 0xFFFF).ne(0).sum()) == 0
    if u < 1:
        assert (st_ref == 1).sum() > 0          # the cap binds somewhere: the case is what it claims to be
    (w * t(G, dtype)).sum().backward()
    dA_ref, db_ref = oracle.risk_budgeting_bwd(G, w_ref, st_ref, tA.detach().double().cpu().numpy(), tb.detach().double().cpu().numpy(), u)
    assert rel_err(tb.grad, db_ref) < 2 * G_RTOL[dtype], rel_err(tb.grad, db_ref)
    assert rel_err(tA.grad, dA_ref) < 2 * G_RTOL[dtype], rel_err(tA.grad, dA_ref)


def test_risk_budgeting_reference_behaviour():
    """tests/test_layers.py:1003-1030 restated: shape, budget, dtype, device, identical problems give identical weights; plus
    the SolverError conventions of the other allocators, zero parameters and pickling."""
    import pickle
    from deepdow_b200 import NumericalRiskBudgeting
    n_samples, n_assets = 5, 6
    for dtype in (torch.float32, torch.float64):
        popt = NumericalRiskBudgeting(n_assets)
        c = torch.rand(n_assets, n_assets, device=DEV, dtype=dtype)
        c = c @ c + torch.eye(n_assets, device=DEV, dtype=dtype)
        covmat_sqrt = torch.stack(n_samples * [c])
        budgets = torch.rand(n_samples, n_assets, device=DEV, dtype=dtype)
        budgets /= budgets.sum(-1, keepdim=True)
        weights = popt(covmat_sqrt, budgets)
        assert weights.shape == (n_samples, n_assets)
        assert torch.allclose(weights.sum(-1), torch.ones(n_samples, device=DEV, dtype=dtype))
        assert weights.dtype == dtype and weights.device == torch.device(DEV)
    assert sum(p.numel() for p in popt.parameters()) == 0
    assert isinstance(pickle.loads(pickle.dumps(popt)), NumericalRiskBudgeting)
    with pytest.raises(SolverError):
        NumericalRiskBudgeting(4, 0.2)(covmat_sqrt[:, :4, :4], budgets[:, :4])
    w = NumericalRiskBudgeting(5, 0.2)(covmat_sqrt[:, :5, :5], budgets[:, :5])
    assert torch.allclose(w, torch.full_like(w, 0.2))
    # equal budgets and a diagonal risk matrix: the closed form w_i ~ 1 / sigma_i (risk parity)
    sig = torch.tensor([1.0, 2.0, 4.0, 8.0], device=DEV, dtype=torch.float64)
    w = NumericalRiskBudgeting(4)(torch.diag(sig)[None], torch.full((1, 4), 0.25, device=DEV, dtype=torch.float64))
    # stationarity: sigma_i^2 w_i - b / w_i + nu = 0 with sum w = 1; compare against the oracle and the identity
    w_ref = oracle.risk_budgeting_fwd(torch.diag(sig)[None].cpu().numpy(), np.full((1, 4), 0.25), 1.0)[0]
    assert np.abs(w.cpu().numpy() - w_ref).max() < 1e-9


# ------------------------------------------------------------------ backward from the factors the forward kept
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("B,N,u,kind", [(64, 100, 0.2, "dense"), (64, 50, 1.0, "covsq"), (96, 100, 0.2, "sym"), (8, 150, 0.05, "dense"),
                                        (4, 240, 0.05, "dense"), (3, 500, 0.01, "sym")])
def test_markowitz_saved_factor_backward_equals_the_regular_one(dtype, B, N, u, kind):
    """ddw_markowitz_fwd_saved / _bwd_saved (what the layer uses when a backward follows) against ddw_markowitz_fwd / _bwd
    called directly through the C ABI: same weights bit for bit, gradients equal to rounding of a different elimination order."""
    from deepdow_b200 import _C
    lib = _C.lib()
    rng = np.random.default_rng(B + N)
    rets, C, g, a = make_problems(rng, B, N, kind, 1.0, 0.5, "rand")
    G = rng.standard_normal((B, N))
    tr, tC, tg, ta, tG = (t(x, dtype) for x in (rets, C, g, a, G))
    dt = _C.dtype_code(tr)
    outs = []
    for use_saved in (False, True):
        w = torch.empty_like(tr); st = torch.empty(B, N, dtype=torch.int8, device=DEV); status = torch.empty(B, dtype=torch.int32, device=DEV)
        ws = torch.empty(lib.ddw_markowitz_workspace_bytes(B, N, dt), dtype=torch.uint8, device=DEV)
        d_r, d_C, d_g, d_a = torch.empty_like(tr), torch.empty_like(tC), torch.empty_like(tg), torch.empty_like(ta)
        if use_saved:
            sv = torch.empty(lib.ddw_markowitz_saved_bytes(B, N, dt), dtype=torch.uint8, device=DEV)
            rc = lib.ddw_markowitz_fwd_saved(_C.ptr(tr), _C.ptr(tC), _C.ptr(tg), _C.ptr(ta), u, B, N, dt, _C.ptr(w), _C.ptr(st), _C.ptr(status),
                                             _C.ptr(ws), ws.numel(), _C.ptr(sv), sv.numel(), _C.stream())
            assert rc == 0, lib.ddw_last_error()
            kept = sv[:4 * B].view(torch.int32)
            rc = lib.ddw_markowitz_bwd_saved(_C.ptr(tG), _C.ptr(w), _C.ptr(st), _C.ptr(tC), _C.ptr(tg), _C.ptr(ta), u, B, N, dt, _C.ptr(d_r),
                                             _C.ptr(d_C), _C.ptr(d_g), _C.ptr(d_a), _C.ptr(ws), ws.numel(), _C.ptr(sv), sv.numel(), _C.stream())
            assert rc == 0, lib.ddw_last_error()
            if kind in ("dense", "covsq") or N > 128:
                assert int((kept >= 0).sum()) > 0          # the dense kernels did keep factors
        else:
            rc = lib.ddw_markowitz_fwd(_C.ptr(tr), _C.ptr(tC), _C.ptr(tg), _C.ptr(ta), u, B, N, dt, _C.ptr(w), _C.ptr(st), _C.ptr(status),
                                       _C.ptr(ws), ws.numel(), _C.stream())
            assert rc == 0, lib.ddw_last_error()
            rc = lib.ddw_markowitz_bwd(_C.ptr(tG), _C.ptr(w), _C.ptr(st), _C.ptr(tC), _C.ptr(tg), _C.ptr(ta), u, B, N, dt, _C.ptr(d_r),
                                       _C.ptr(d_C), _C.ptr(d_g), _C.ptr(d_a), _C.ptr(ws), ws.numel(), _C.stream())
            assert rc == 0, lib.ddw_last_error()
        torch.cuda.synchronize()
        outs.append((w, st, d_r, d_C, d_g, d_a))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    tol = 2e-4 if dtype == torch.float32 else 1e-9
    for name, x, y in zip(["rets", "covmat_sqrt", "gamma", "alpha"], outs[0][2:], outs[1][2:]):
        assert rel_err(y, x.double().cpu().numpy()) < tol, (name, rel_err(y, x.double().cpu().numpy()))


# ------------------------------------------------------------------ robustness
def test_markowitz_failure_in_the_last_batch_is_reported():
    """Lazy status + a Run.launch-style loop (experiments.py:322-390): the batch with a non-finite input is the LAST one of
    the epoch, nothing calls the layer again -- the switch to eval() (validation, callbacks.py:814) must raise SolverError.
    With check_status=True the failing call itself raises."""
    rng = np.random.default_rng(8)
    rets, C, g, a = make_problems(rng, 16, 12)
    good = [t(x, torch.float32) for x in (rets, C, g, a)]
    bad_rets = rets.copy(); bad_rets[3, 5] = np.nan
    layer = NumericalMarkowitz(12, 0.5)
    assert layer.check_status == "lazy"
    for _ in range(3):
        layer(*good)
    w = layer(t(bad_rets, torch.float32), *good[1:])             # last batch of the "epoch"
    with pytest.raises(SolverError):
        layer.eval()
    layer.train()                                                # the report was consumed: the next epoch starts clean
    layer(*good); layer.synchronize_status()
    strict = NumericalMarkowitz(12, 0.5); strict.check_status = True
    with pytest.raises(SolverError):
        strict(t(bad_rets, torch.float32), *good[1:])
    bad_C = C.copy(); bad_C[2, 1, 1] = np.inf
    with pytest.raises(SolverError):
        strict(good[0], t(bad_C, torch.float32), *good[2:])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_layers_on_a_second_gpu_in_the_same_process():
    """The > 48 KB shared-memory opt-in belongs to a device: every layer must work on cuda:1 after cuda:0 in one process."""
    from deepdow_b200 import NumericalRiskBudgeting
    rng = np.random.default_rng(9)
    rets, C, g, a = make_problems(rng, 8, 100)
    x = rng.standard_normal((4, 40, 50)) * 0.01
    xl = rng.standard_normal((2, 40, 200)) * 0.01
    outs = []
    for d in ("cuda:0", "cuda:1"):
        ins = [torch.tensor(v, dtype=torch.float32, device=d, requires_grad=True) for v in (rets, C, g, a)]
        w = NumericalMarkowitz(100, 0.2)(*ins)
        grads = torch.autograd.grad(w, ins, torch.ones_like(w))
        dense = NumericalMarkowitz(100, 0.2)(ins[0].detach() * 0.04, *[v.detach() for v in ins[1:]])
        cov = CovarianceMatrix(sqrt=True)(torch.tensor(x, dtype=torch.float32, device=d))
        covp = CovarianceMatrix(sqrt=False)(torch.tensor(x, dtype=torch.float32, device=d))
        covl = CovarianceMatrix(sqrt=True)(torch.tensor(xl, dtype=torch.float32, device=d))
        sm = SparsemaxAllocator(100, 1.0, 0.05)(ins[0].detach())
        b = torch.full((8, 100), 0.01, device=d)
        rb = NumericalRiskBudgeting(100)(ins[1].detach(), b)
        outs.append([v.cpu() for v in (w.detach(), grads[1], dense, cov, covp, covl, sm, rb)])
    for x0, x1 in zip(*outs):
        assert torch.equal(x0, x1)


# ------------------------------------------------------------------ fused covariance -> NumericalMarkowitz
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("B,L,N,u,strategy,rscale", [(64, 40, 50, 1.0, "diagonal", 0.01), (32, 40, 100, 0.2, "diagonal", 0.01),
                                                      (16, 12, 20, 1.0, "scaled_identity", 0.01), (16, 30, 37, 0.5, None, 0.01),
                                                      (32, 40, 50, 1.0, "identity", 2.0)])
def test_fused_covariance_markowitz_equals_the_composition(dtype, B, L, N, u, strategy, rscale):
    """ddw_covariance_markowitz_fwd against CovarianceMatrix(sqrt=False) -> NumericalMarkowitz: in the diversified regime both
    run the dense solver on the same bits (weights bit for bit); with concentrated optima or small universes the composition
    takes the sparse-support / warp-per-portfolio kernels, so weights agree to rounding.  Gradients flow to x, coef, rets, gamma and alpha; oracle check."""
    from deepdow_b200.layers import covariance_markowitz
    rng = np.random.default_rng(B + L + N)
    x = rng.standard_normal((B, L, N)) * 0.01
    rets = rng.standard_normal((B, N)) * rscale
    G = rng.standard_normal((B, N))
    coef = rng.uniform(0.3, 0.8, B)
    res = []
    for fused in (False, True):
        tx, tr = t(x, dtype, True), t(rets, dtype, True)
        tc = t(coef, dtype, True) if strategy is not None else None
        tg, ta = t(np.full(B, 1.5), dtype, True), t(np.full(B, 0.7), dtype, True)
        cov_layer = CovarianceMatrix(sqrt=False, shrinkage_strategy=strategy, shrinkage_coef=None if strategy else 0.5)
        alloc = NumericalMarkowitz(N, u)
        if fused:
            w = covariance_markowitz(tx, tr, tg, ta, cov_layer, alloc, tc)
        else:
            w = alloc(tr, cov_layer(tx, tc), tg, ta)
        (w * t(G, dtype)).sum().backward()
        res.append((w.detach(), tx.grad, tr.grad, tg.grad, ta.grad, None if tc is None else tc.grad))
    # float32 with at most 32 + 8 free assets: the composition runs the warp-per-portfolio solver on the tensor-core Gram
    # (markowitz_qwarp.cu), the fused call the dense solver -- equal to rounding, not bit for bit
    dense = rscale < 1.0 and (dtype == torch.float64 or N > 40)
    if dense:
        assert torch.equal(res[0][0], res[1][0])
    else:
        assert torch.allclose(res[0][0], res[1][0], atol=2e-6 if dtype == torch.float32 else 1e-12)
    for a_, b_ in zip(res[0][1:], res[1][1:]):
        if a_ is not None:
            assert rel_err(b_, a_.double().cpu().numpy()) < (2e-4 if dtype == torch.float32 else 1e-9)
    C = oracle.covariance_fwd(t(x, dtype).double().cpu().numpy(), coef if strategy else 0.5, strategy, False)
    w_ref, *_ = oracle.markowitz_fwd(t(rets, dtype).double().cpu().numpy(), C, np.full(B, 1.5), np.full(B, 0.7), u)
    assert np.abs(res[1][0].double().cpu().numpy() - w_ref).max() < W_ATOL[dtype]

"""Pin the CPU oracle: reference golden vectors, fixtures produced by the reference's own
misc.py, scipy SLSQP fixtures, KKT certificates and finite differences.  CPU only."""
import os

import numpy as np
import pytest

import oracle


@pytest.fixture(scope="module")
def refvec(golden_dir):
    return np.load(os.path.join(golden_dir, "reference_vectors.npz"))


# ---------------------------------------------------------------- sparsemax
def test_sparsemax_known(refvec):
    # reference tests/test_layers.py:821-838
    w = oracle.capped_simplex_fwd(refvec["sparsemax_known_x"], 1.0, 1.0)
    assert np.allclose(w, refvec["sparsemax_known_w"], atol=1e-3)
    assert np.allclose(w.sum(1), 1.0, atol=1e-12)


def test_sparsemax_doc(refvec):
    # docs/source/layers.rst:381-394
    w = oracle.capped_simplex_fwd(refvec["sparsemax_doc_x"], 1.0, 1.0)
    assert np.allclose(w, refvec["sparsemax_doc_w"], atol=2e-4)
    assert np.allclose(w, [[0, 0.6, 0.4], [0, 1, 0]], atol=1e-12)


def test_sparsemax_capped(refvec):
    # tests/test_layers.py:840-850 and SURVEY 8(c) exact values
    exact = {0.2: [0.2] * 5, 0.25: [0.25, 0, 0.25, 0.25, 0.25], 0.34: [0.34, 0, 0.0677, 0.2523, 0.34]}
    for u in refvec["capped_u"]:
        w = oracle.capped_simplex_fwd(refvec["capped_x"], 1.0, float(u))
        assert w.max() == pytest.approx(u, abs=1e-12)
        assert np.allclose(w[0], exact[float(u)], atol=1e-4)
        assert w.sum() == pytest.approx(1.0, abs=1e-12)


def test_sparsemax_uniform_and_temperature():
    w = oracle.capped_simplex_fwd(np.ones((2, 5)), 1.0)
    assert np.allclose(w, 0.2)
    rng = np.random.default_rng(0)
    x = rng.standard_normal((4, 7))
    assert np.allclose(oracle.capped_simplex_fwd(x, 2.0), oracle.capped_simplex_fwd(x / 2.0, np.ones(4)))


def _fd(fun, x, eps=1e-6):
    g = np.zeros_like(x)
    it = np.nditer(x, flags=["multi_index"])
    for _ in it:
        i = it.multi_index
        xp = x.copy(); xp[i] += eps
        xm = x.copy(); xm[i] -= eps
        g[i] = (fun(xp) - fun(xm)) / (2 * eps)
    return g


def test_sparsemax_bwd_fd():
    rng = np.random.default_rng(1)
    x = rng.standard_normal((3, 9)) * 0.5
    t = np.array([0.7, 1.0, 1.9])
    G = rng.standard_normal((3, 9))
    for u in [1.0, 0.3]:
        w = oracle.capped_simplex_fwd(x, t, u)
        dx, dt = oracle.capped_simplex_bwd(G, w, x, t, u)
        f = lambda xx: (oracle.capped_simplex_fwd(xx, t, u) * G).sum()
        assert np.allclose(dx, _fd(f, x), atol=1e-6)
        ft = lambda tt: (oracle.capped_simplex_fwd(x, tt, u) * G).sum()
        assert np.allclose(dt, _fd(ft, t), atol=1e-6)


# ---------------------------------------------------------------- softmax
def test_softmax_equals_analytical():
    # tests/test_layers.py:726-750
    rng = np.random.default_rng(2)
    x = rng.integers(0, 10, (3, 6)) + rng.random((3, 6))
    e = np.exp(x - x.max(1, keepdims=True))
    assert np.allclose(oracle.capped_softmax_fwd(x, 1.0, 1.0), e / e.sum(1, keepdims=True), atol=1e-12)


def test_softmax_capped(refvec):
    # tests/test_layers.py:761-776 + SURVEY 8(c) values
    exact = {0.25: [0.25, 0.0542, 0.2024, 0.2434, 0.25], 0.34: [0.34, 0.0391, 0.1462, 0.1758, 0.2989]}
    for u in refvec["capped_u"]:
        w = oracle.capped_softmax_fwd(refvec["capped_x"], 1.0, float(u))
        assert w.max() == pytest.approx(u, abs=1e-12)
        assert w.sum() == pytest.approx(1.0, abs=1e-12)
        if float(u) in exact:
            assert np.allclose(w[0], exact[float(u)], atol=1e-4)


def test_softmax_bwd_fd():
    rng = np.random.default_rng(3)
    x = rng.standard_normal((3, 8))
    t = np.array([0.5, 1.0, 2.0])
    G = rng.standard_normal((3, 8))
    for u in [1.0, 0.2]:
        w = oracle.capped_softmax_fwd(x, t, u)
        dx, dt = oracle.capped_softmax_bwd(G, w, x, t, u)
        f = lambda xx: (oracle.capped_softmax_fwd(xx, t, u) * G).sum()
        assert np.allclose(dx, _fd(f, x), atol=1e-6)
        ft = lambda tt: (oracle.capped_softmax_fwd(x, tt, u) * G).sum()
        assert np.allclose(dt, _fd(ft, t), atol=1e-6)


# ---------------------------------------------------------------- covariance
def test_covariance_vs_reference_fixtures(golden_dir):
    z = np.load(os.path.join(golden_dir, "covariance_reference.npz"))
    for c in range(int(z["n_cases"])):
        k = f"c{c}"
        strategy, sqrt = z[k + "_meta"]
        strategy = None if strategy == "None" else str(strategy)
        sqrt = bool(int(sqrt))
        y = oracle.covariance_fwd(z[k + "_x"], z[k + "_coef"], strategy, sqrt)
        scale = np.abs(z[k + "_y"]).max()
        assert np.allclose(y, z[k + "_y"], atol=1e-10 * scale), (c, strategy, sqrt)
        dx, dc = oracle.covariance_bwd(z[k + "_G"], z[k + "_x"], z[k + "_coef"], strategy, sqrt)
        if not np.isfinite(z[k + "_dx"]).all():
            # identity-type shrinkage with L - 1 < N gives repeated eigenvalues: the reference's SVD
            # autograd (misc.py:183) returns NaN there.  The Sylvester form stays finite; check it
            # against finite differences instead.
            assert sqrt and strategy in ("identity", "scaled_identity")
            assert np.isfinite(dx).all() and np.isfinite(dc).all()
            f = lambda xx: (oracle.covariance_fwd(xx, z[k + "_coef"], strategy, sqrt) * z[k + "_G"]).sum()
            x0 = z[k + "_x"]
            h = 1e-6 * np.abs(x0).max()
            for idx in [(0, 0, 0), (1, 3, 2), (0, 5, 7)]:
                xp = x0.copy(); xp[idx] += h
                xm = x0.copy(); xm[idx] -= h
                fd = (f(xp) - f(xm)) / (2 * h)
                assert fd == pytest.approx(dx[idx], rel=1e-4, abs=1e-6 * np.abs(dx).max())
            continue
        assert np.allclose(dx, z[k + "_dx"], rtol=1e-6, atol=1e-8 * np.abs(z[k + "_dx"]).max()), (c, strategy, sqrt)
        assert np.allclose(dc, z[k + "_dcoef"], rtol=1e-6, atol=1e-8 * max(1e-30, np.abs(z[k + "_dcoef"]).max())), (c, strategy, sqrt)


def test_covariance_rank_deficient_and_doc(golden_dir):
    z = np.load(os.path.join(golden_dir, "covariance_reference.npz"))
    y = oracle.covariance_fwd(z["rd_x"], 1.0, None, True)
    assert np.allclose(y, z["rd_y"], atol=1e-9)
    cov = oracle.covariance_fwd(z["doc_x"], 0.5, "diagonal", False)
    sq = oracle.covariance_fwd(z["doc_x"], 0.5, "diagonal", True)
    assert np.allclose(cov, z["doc_cov"], rtol=1e-4)
    assert np.allclose(sq, z["doc_sqrt"], rtol=1e-4, atol=1e-3)
    assert np.allclose(sq[0] @ sq[0], cov[0], atol=1e-8 * cov.max())


def test_covariance_dim1_raises():
    with pytest.raises(ZeroDivisionError):
        oracle.covariance_fwd(np.ones((2, 1, 3)), 0.5)


# ---------------------------------------------------------------- markowitz
def _problems(rng, B, N, kind="sym", alpha=1.0, rscale=0.5):
    rets = rng.standard_normal((B, N)) * rscale
    if kind == "sym":
        R = rng.standard_normal((B, N, N))
        C = R @ R.transpose(0, 2, 1) / N + 0.1 * np.eye(N)
    else:
        C = rng.standard_normal((B, N, N)) * 0.5 + np.eye(N)
    return rets, C, np.ones(B), np.full(B, alpha)


def test_markowitz_kkt_certificate():
    rng = np.random.default_rng(5)
    for N, u in [(5, 1.0), (20, 0.2), (50, 0.05), (100, 0.2), (100, 1.0)]:
        rets, C, g, a = _problems(rng, 16, N)
        w, st, kkt, status = oracle.markowitz_fwd(rets, C, g, a, u)
        assert (status == 0).all()
        assert kkt.max() < 1e-11
        assert np.allclose(w.sum(1), 1.0, atol=1e-12)
        assert w.min() >= 0 and w.max() <= u + 1e-15


def test_markowitz_vs_slsqp_fixture(golden_dir):
    z = np.load(os.path.join(golden_dir, "markowitz_slsqp.npz"))
    for c in range(int(z["n_cases"])):
        k = f"m{c}"
        w, st, kkt, status = oracle.markowitz_fwd(z[k + "_rets"], z[k + "_C"], z[k + "_gamma"], z[k + "_alpha"], float(z[k + "_u"]))
        assert (status == 0).all() and kkt.max() < 1e-11
        assert np.allclose(w, z[k + "_w"], atol=2e-6), (c, np.abs(w - z[k + "_w"]).max())


def test_markowitz_reduces_to_sparsemax(refvec):
    # covmat_sqrt = I, alpha = 0, gamma = 1: min w.w - r.w == projection of r/2
    for x in (refvec["sparsemax_known_x"], refvec["sparsemax_doc_x"]):
        B, N = x.shape
        C = np.broadcast_to(np.eye(N), (B, N, N)).copy()
        w, _, kkt, _ = oracle.markowitz_fwd(2 * x, C, np.ones(B), np.zeros(B), 1.0)
        assert np.allclose(w, oracle.capped_simplex_fwd(x, 1.0), atol=1e-12)
    w, _, _, _ = oracle.markowitz_fwd(2 * refvec["sparsemax_known_x"], np.broadcast_to(np.eye(5), (2, 5, 5)).copy(),
                                      np.ones(2), np.zeros(2), 1.0)
    assert np.allclose(w, refvec["sparsemax_known_w"], atol=1e-3)
    for u in refvec["capped_u"]:
        w, _, _, _ = oracle.markowitz_fwd(2 * refvec["capped_x"], np.eye(5)[None], np.ones(1), np.zeros(1), float(u))
        assert np.allclose(w, oracle.capped_simplex_fwd(refvec["capped_x"], 1.0, float(u)), atol=1e-12)


def test_markowitz_vertex_and_infeasible():
    rng = np.random.default_rng(6)
    # LP-like: tiny risk, u = 0.2 -> exactly five assets at the cap
    rets = rng.standard_normal((4, 30))
    C = np.broadcast_to(np.eye(30) * 1e-3, (4, 30, 30)).copy()
    w, st, kkt, status = oracle.markowitz_fwd(rets, C, np.ones(4), np.zeros(4), 0.2)
    assert (status == 0).all() and kkt.max() < 1e-10
    for b in range(4):
        top = np.argsort(-rets[b])[:5]
        assert np.allclose(w[b, top], 0.2, atol=1e-9)
    # N * u == 1: single feasible point
    w, st, kkt, status = oracle.markowitz_fwd(rets[:, :5], C[:, :5, :5], np.ones(4), np.ones(4), 0.2)
    assert np.allclose(w, 0.2) and (status == 0).all()
    # N * u < 1: infeasible
    _, _, _, status = oracle.markowitz_fwd(rets[:, :4], C[:, :4, :4], np.ones(4), np.ones(4), 0.2)
    assert (status == 1).all()


def test_markowitz_gamma_tiling():
    # allocate.py:268-270: repeat/view is a tiling, not a per-sample broadcast
    import torch
    rng = np.random.default_rng(7)
    B, N = 5, 4
    rets, C, _, a = _problems(rng, B, N)
    gamma = rng.uniform(0.5, 2.0, B)
    gamma_ = torch.from_numpy(gamma).repeat((1, N * N)).view(B, N, N).numpy()
    w1, *_ = oracle.markowitz_fwd(rets, C, gamma, a, 1.0)
    w2, *_ = oracle.markowitz_fwd(rets, gamma_ * C, np.ones(B), a, 1.0)
    assert np.allclose(w1, w2, atol=1e-13)


def test_markowitz_bwd_fd():
    rng = np.random.default_rng(8)
    B, N, u = 3, 7, 0.35
    rets, C, _, _ = _problems(rng, B, N, kind="asym")
    gamma = rng.uniform(0.7, 1.5, B)
    alpha = np.array([0.8, -1.2, 0.5])
    G = rng.standard_normal((B, N))
    w, st, kkt, _ = oracle.markowitz_fwd(rets, C, gamma, alpha, u)
    assert ((st != 0).sum() > 0) and ((st == 0).sum() > B)  # some active, some free
    d_r, d_C, d_g, d_a = oracle.markowitz_bwd(G, w, st, C, gamma, alpha, u)
    L = lambda r_, C_, g_, a_: (oracle.markowitz_fwd(r_, C_, g_, a_, u)[0] * G).sum()
    assert np.allclose(d_r, _fd(lambda v: L(v, C, gamma, alpha), rets), atol=2e-6)
    assert np.allclose(d_C, _fd(lambda v: L(rets, v, gamma, alpha), C), atol=2e-6)
    assert np.allclose(d_g, _fd(lambda v: L(rets, C, v, alpha), gamma), atol=2e-6)
    assert np.allclose(d_a, _fd(lambda v: L(rets, C, gamma, v), alpha), atol=2e-6)


# ------------------------------------------------------------------ NumericalRiskBudgeting oracle (allocate.py:651-691)
def test_risk_budgeting_oracle_kkt_scipy_and_finite_differences():
    """The reference holds no numeric golden for NumericalRiskBudgeting (tests/test_layers.py:1003-1030 checks shape, budget
    and dtype only): the oracle is anchored by its KKT certificate, by scipy SLSQP on the same program and by central finite
    differences of both gradients; with a loose cap the risk contributions w_i (Q w)_i equal the budgets up to the multiplier."""
    import scipy.optimize as so
    rng = np.random.default_rng(0)
    for (B, N, u) in [(3, 6, 1.0), (2, 10, 0.2), (2, 30, 0.05)]:
        R = rng.random((B, N, N))
        A = R @ R + np.eye(N)                                   # as tests/test_layers.py:1011-1015
        b = rng.random((B, N)); b /= b.sum(1, keepdims=True)
        w, st, kkt, status = oracle.risk_budgeting_fwd(A, b, u)
        assert kkt.max() < 1e-9 and (status == 0).all()
        assert np.allclose(w.sum(1), 1.0, atol=1e-12) and w.min() > 0 and w.max() <= u + 1e-15
        for i in range(B):
            Q = A[i].T @ A[i]
            res = so.minimize(lambda x: 0.5 * x @ Q @ x - b[i] @ np.log(x), np.full(N, 1.0 / N), jac=lambda x: Q @ x - b[i] / x,
                              method="SLSQP", bounds=[(1e-9, u)] * N,
                              constraints=[{"type": "eq", "fun": lambda x: x.sum() - 1, "jac": lambda x: np.ones(N)}],
                              options={"ftol": 1e-16, "maxiter": 2000})
            assert np.abs(res.x - w[i]).max() < 1e-6
        if u == 1.0:                                            # risk budgeting identity: w_i (Q w)_i - b_i = -nu w_i
            Qw = np.einsum("bji,bjk,bk->bi", A, A, w)
            nu = (b - w * Qw) / w
            assert np.abs(nu - nu[:, :1]).max() < 1e-8
        G = rng.standard_normal((B, N))
        dA, db = oracle.risk_budgeting_bwd(G, w, st, A, b, u)
        eps = 1e-6
        for j in range(N):
            bp, bm = b[:1].copy(), b[:1].copy()
            bp[0, j] += eps; bm[0, j] -= eps
            fd = G[0] @ (oracle.risk_budgeting_fwd(A[:1], bp, u)[0][0] - oracle.risk_budgeting_fwd(A[:1], bm, u)[0][0]) / (2 * eps)
            assert abs(fd - db[0, j]) < 1e-5 * max(1.0, np.abs(db[0]).max())
        for (j, k) in [(0, 0), (1, 2), (N - 1, 0)]:
            Ap, Am = A[:1].copy(), A[:1].copy()
            Ap[0, j, k] += eps; Am[0, j, k] -= eps
            fd = G[0] @ (oracle.risk_budgeting_fwd(Ap, b[:1], u)[0][0] - oracle.risk_budgeting_fwd(Am, b[:1], u)[0][0]) / (2 * eps)
            assert abs(fd - dA[0, j, k]) < 1e-4 * max(1e-3, np.abs(dA[0]).max())
    # box-only cases
    w, st, kkt, status = oracle.risk_budgeting_fwd(np.eye(4)[None], np.full((1, 4), 0.25), 0.25)
    assert np.allclose(w, 0.25) and (st == 1).all()
    assert oracle.risk_budgeting_fwd(np.eye(4)[None], np.full((1, 4), 0.25), 0.2)[3][0] == 8

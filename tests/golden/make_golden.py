"""Generate the golden fixtures in this directory.  Run in the DEV container only:

    python tests/golden/make_golden.py

* covariance_*.npz -- produced by the UNMODIFIED reference ``deepdow/layers/misc.py`` loaded by file
  path from /root/reference (it imports only torch); forward values and autograd gradients, float64.
* markowitz_slsqp.npz -- the reference holds no numeric golden for NumericalMarkowitz and
  cvxpylayers is not installable here, so the program stated at allocate.py:220-232 (with the
  forward pre-processing of allocate.py:267-273 done by the same torch calls) is solved by
  scipy's SLSQP, an independent third-party solver, in float64.
* reference_vectors.npz -- golden vectors copied from the reference's tests and docs
  (tests/test_layers.py:821-838, docs/source/layers.rst:381-394).

/root/reference does not exist on the GPU box; only the .npz files travel.
"""
import importlib.util
import os

import numpy as np
import scipy.optimize as so
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/deepdow/layers/misc.py"


def load_ref_misc():
    spec = importlib.util.spec_from_file_location("ref_misc", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def covariance():
    misc = load_ref_misc()
    out = {}
    case = 0
    for (B, L, N, scale) in [(3, 7, 5, 1.0), (2, 40, 20, 0.01), (2, 12, 16, 100.0)]:
        for strategy in ["diagonal", "identity", "scaled_identity", None]:
            for sqrt in [True, False]:
                if strategy is None and sqrt and L - 1 < N:
                    continue  # rank deficient: SVD autograd is ill-defined; forward-only case below
                g = torch.Generator().manual_seed(100 + case)
                x = (torch.randn(B, L, N, generator=g, dtype=torch.float64) * scale).requires_grad_()
                coef = (0.2 + 0.6 * torch.rand(B, generator=g, dtype=torch.float64)).requires_grad_()
                layer = misc.CovarianceMatrix(sqrt=sqrt, shrinkage_strategy=strategy, shrinkage_coef=None)
                y = layer(x, coef)
                G = torch.randn(y.shape, generator=g, dtype=torch.float64)
                y.backward(G)
                k = f"c{case}"
                out[k + "_x"] = x.detach().numpy()
                out[k + "_coef"] = coef.detach().numpy()
                out[k + "_y"] = y.detach().numpy()
                out[k + "_G"] = G.numpy()
                out[k + "_dx"] = x.grad.numpy()
                out[k + "_dcoef"] = (coef.grad.numpy() if coef.grad is not None else np.zeros(B))
                out[k + "_meta"] = np.array([str(strategy), str(int(sqrt))])
                case += 1
    out["n_cases"] = np.array(case)
    # rank-deficient forward (strategy None, L - 1 < N), float64 and float32
    g = torch.Generator().manual_seed(7)
    x = torch.randn(2, 4, 6, generator=g, dtype=torch.float64)
    out["rd_x"] = x.numpy()
    out["rd_y"] = misc.CovarianceMatrix(sqrt=True, shrinkage_strategy=None, shrinkage_coef=1.0)(x).numpy()
    # docs/source/layers.rst:474-487 doctest inputs (float32)
    torch.manual_seed(3)
    xd = torch.rand(1, 10, 3) * 100
    out["doc_x"] = xd.numpy()
    out["doc_cov"] = misc.CovarianceMatrix(sqrt=False)(xd).numpy()
    out["doc_sqrt"] = misc.CovarianceMatrix(sqrt=True)(xd).numpy()
    np.savez_compressed(os.path.join(HERE, "covariance_reference.npz"), **out)
    print("covariance cases:", case)


def markowitz():
    rng = np.random.default_rng(2024)
    out = {}
    case = 0
    for (B, N, u, alpha_scale, kind) in [
        (6, 5, 1.0, 1.0, "sym"), (6, 8, 0.3, 1.0, "sym"), (5, 12, 0.2, 0.5, "asym"),
        (4, 10, 0.25, 0.0, "sym"), (7, 6, 1.0, 1.0, "asym"), (3, 20, 0.1, 1.0, "sym"),
    ]:
        rets = rng.standard_normal((B, N)) * 0.5
        if kind == "sym":
            R = rng.standard_normal((B, N, N))
            C = R @ R.transpose(0, 2, 1) / N + 0.1 * np.eye(N)
        else:
            C = rng.standard_normal((B, N, N)) * 0.5 + np.eye(N)
        gamma = rng.uniform(0.5, 2.0, B)
        alpha = rng.uniform(0.5, 1.5, B) * alpha_scale * rng.choice([-1.0, 1.0], B)
        # forward pre-processing exactly as allocate.py:267-273 (torch API calls)
        g_t = torch.from_numpy(gamma)
        gamma_ = g_t.repeat((1, N * N)).view(B, N, N).numpy()
        A = gamma_ * C
        W = np.empty((B, N))
        for b in range(B):
            Ab, rb, ab = A[b], rets[b], abs(alpha[b])
            f = lambda w: -(rb @ w) + np.sum((Ab @ w) ** 2) + ab * (w @ w)
            df = lambda w: -rb + 2 * Ab.T @ (Ab @ w) + 2 * ab * w
            best = None
            for x0 in [np.full(N, 1.0 / N), rng.dirichlet(np.ones(N))]:
                x0 = np.clip(x0, 0, u); x0 = x0 / x0.sum()
                res = so.minimize(f, x0, jac=df, method="SLSQP", bounds=[(0, u)] * N,
                                  constraints=[{"type": "eq", "fun": lambda w: w.sum() - 1,
                                                "jac": lambda w: np.ones(N)}],
                                  options={"ftol": 1e-15, "maxiter": 1000})
                if best is None or res.fun < best.fun:
                    best = res
            W[b] = best.x
        k = f"m{case}"
        out[k + "_rets"] = rets; out[k + "_C"] = C; out[k + "_gamma"] = gamma; out[k + "_alpha"] = alpha
        out[k + "_u"] = np.array(u); out[k + "_w"] = W
        case += 1
    out["n_cases"] = np.array(case)
    np.savez_compressed(os.path.join(HERE, "markowitz_slsqp.npz"), **out)
    print("markowitz cases:", case)


def markowitz_large():
    """Independent third-party fixtures at n_assets = 50 and 100 (scipy SLSQP, restarted once), including the diversified
    ("dense": most assets free) and the cap-bound regimes of BASELINE configs 2 / 4.  Written to their own file so that
    markowitz_slsqp.npz keeps its bytes."""
    rng = np.random.default_rng(4048)
    out = {}
    case = 0
    for (B, N, u, rscale, gam) in [(3, 50, 1.0, 0.5, "rand"), (3, 50, 0.1, 0.02, "ones"), (2, 100, 0.2, 0.5, "ones"),
                                   (2, 100, 0.2, 0.02, "ones"), (2, 100, 0.05, 3.0, "rand")]:
        rets = rng.standard_normal((B, N)) * rscale
        R = rng.standard_normal((B, N, N))
        C = R @ R.transpose(0, 2, 1) / N + 0.1 * np.eye(N)
        gamma = rng.uniform(0.5, 2.0, B) if gam == "rand" else np.ones(B)
        alpha = rng.uniform(0.5, 1.5, B) * rng.choice([-1.0, 1.0], B)
        gamma_ = torch.from_numpy(gamma).repeat((1, N * N)).view(B, N, N).numpy()     # allocate.py:268-270
        A = gamma_ * C
        W = np.empty((B, N))
        for b in range(B):
            Ab, rb, ab = A[b], rets[b], abs(alpha[b])
            H = 2 * (Ab.T @ Ab + ab * np.eye(N))
            f = lambda w: -(rb @ w) + 0.5 * w @ H @ w
            df = lambda w: -rb + H @ w
            x0 = np.clip(np.full(N, 1.0 / N), 0, u); x0 /= x0.sum()
            res = so.minimize(f, x0, jac=df, method="SLSQP", bounds=[(0, u)] * N,
                              constraints=[{"type": "eq", "fun": lambda w: w.sum() - 1, "jac": lambda w: np.ones(N)}],
                              options={"ftol": 1e-16, "maxiter": 2000})
            # restart from the first answer: SLSQP's active-set guess is then right and the final QP step is exact to rounding
            res2 = so.minimize(f, res.x, jac=df, method="SLSQP", bounds=[(0, u)] * N,
                               constraints=[{"type": "eq", "fun": lambda w: w.sum() - 1, "jac": lambda w: np.ones(N)}],
                               options={"ftol": 1e-18, "maxiter": 2000})
            W[b] = res2.x if res2.fun <= res.fun else res.x
        k = f"m{case}"
        out[k + "_rets"] = rets; out[k + "_C"] = C; out[k + "_gamma"] = gamma; out[k + "_alpha"] = alpha
        out[k + "_u"] = np.array(u); out[k + "_w"] = W
        case += 1
    out["n_cases"] = np.array(case)
    np.savez_compressed(os.path.join(HERE, "markowitz_scipy_large.npz"), **out)
    print("markowitz large cases:", case)


def reference_vectors():
    out = dict(
        # tests/test_layers.py:821-838 (atol 1e-3)
        sparsemax_known_x=np.array([[1.7909, 0.3637, -0.6818, -0.4972, 0.0333],
                                    [0.6655, -0.9960, 1.1463, 1.9849, -0.1662]]),
        sparsemax_known_w=np.array([[1.0, 0.0, 0.0, 0.0, 0.0], [0.0, 0.0, 0.0807, 0.9193, 0.0]]),
        # docs/source/layers.rst:381-394 (atol 1e-4; the doc output carries SCS noise)
        sparsemax_doc_x=np.array([[1, 2.3, 2.1], [2, 4.2, -1.1]]),
        sparsemax_doc_w=np.array([[-4.8035e-06, 6.0000e-01, 4.0000e-01],
                                  [8.9401e-05, 1.0001e+00, -1.7873e-04]]),
        # tests/test_layers.py:761-776, 840-850: capped logits, w.max() == max_weight
        capped_x=np.array([[1.7909, -2, -0.6818, -0.4972, 0.0333]]),
        capped_u=np.array([0.2, 0.25, 0.34]),
    )
    np.savez_compressed(os.path.join(HERE, "reference_vectors.npz"), **out)


if __name__ == "__main__":
    import sys
    if "--large-only" in sys.argv:
        markowitz_large()
    else:
        covariance()
        markowitz()
        markowitz_large()
        reference_vectors()

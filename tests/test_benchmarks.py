"""MinimumVariance / MaximumReturn (deepdow/benchmarks.py:76-262) on the CUDA path.

The reference checks these only by property (tests/test_benchmarks.py:60-101: shape, dtype, device, sum = 1, bounds), and its
solver is not installable here; parity is therefore against the oracle: the LP's vertex solution (oracle.capped_argmax_fwd,
itself pinned below by brute force and by scipy's linprog) and the Markowitz oracle with rets = 0, alpha = 0.
"""
import itertools

import numpy as np
import pytest
import torch

import oracle


def test_lp_oracle_against_linprog_and_known_answers():
    from scipy.optimize import linprog
    rng = np.random.default_rng(5)
    for N, u in [(5, 1.0), (5, 0.34), (7, 0.25), (12, 0.2), (9, 0.15)]:
        x = rng.standard_normal((6, N))
        w = oracle.capped_argmax_fwd(x, u)
        assert np.allclose(w.sum(1), 1.0) and (w >= 0).all() and (w <= u + 1e-12).all()
        for b in range(len(x)):
            res = linprog(-x[b], A_eq=np.ones((1, N)), b_eq=[1.0], bounds=[(0, u)] * N, method="highs")
            assert res.status == 0
            assert abs(-res.fun - x[b] @ w[b]) < 1e-9               # same optimal value
            assert np.abs(res.x - w[b]).max() < 1e-7                # distinct entries: the optimum is unique
    # max_weight = 1: one-hot on the arg-max; ties go to the smaller index
    w = oracle.capped_argmax_fwd(np.array([[0.1, 0.7, 0.7, -1.0]]), 1.0)
    assert w.tolist() == [[0.0, 1.0, 0.0, 0.0]]
    w = oracle.capped_argmax_fwd(np.array([[3.0, 1.0, 2.0, 0.0]]), 0.4)
    assert np.allclose(w, [[0.4, 0.2, 0.4, 0.0]])
    with pytest.raises(ValueError):
        oracle.capped_argmax_fwd(np.zeros((1, 3)), 0.3)


def test_benchmark_host_logic():
    from deepdow_b200.benchmarks import MaximumReturn, MinimumVariance
    from deepdow_b200._C import ExtensionError
    for cls in (MaximumReturn, MinimumVariance):
        b = cls(max_weight=0.5, n_assets=4, returns_channel=1)
        assert b.hparams == {"max_weight": 0.5, "returns_channel": 1, "n_assets": 4}
        with pytest.raises(ValueError):                 # benchmarks.py:141-146, :236-241
            b(torch.zeros(2, 2, 5, 3))
        with pytest.raises((ExtensionError, ValueError, RuntimeError)):      # no CPU fallback
            cls(n_assets=None)(torch.zeros(2, 1, 5, 3))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("B,L,N,u", [(64, 10, 20, 1.0), (37, 5, 50, 0.1), (9, 8, 100, 0.05), (5, 4, 130, 0.3), (3, 3, 700, 0.01),
                                     (16, 6, 5, 0.34), (2, 4, 2048, 0.001)])
def test_maximum_return_vs_oracle(dtype, B, L, N, u):
    from deepdow_b200.benchmarks import MaximumReturn
    rng = np.random.default_rng(N)
    x = torch.tensor(rng.standard_normal((B, 2, L, N)), dtype=dtype, device="cuda:0")
    w = MaximumReturn(max_weight=u, n_assets=N, returns_channel=1)(x)
    assert w.shape == (B, N) and w.dtype == dtype and w.device == x.device
    rets = x[:, 1].mean(dim=1).double().cpu().numpy()         # the exact vector the kernel ranks
    ref = oracle.capped_argmax_fwd(rets, float(np.float32(u)) if dtype == torch.float32 else u)
    assert np.abs(w.double().cpu().numpy() - ref).max() < (1e-6 if dtype == torch.float32 else 1e-12)
    assert ((w > 0).cpu().numpy() == (ref > 0)).all()         # same funded assets


@pytest.mark.gpu
def test_maximum_return_ties_and_infeasible():
    from deepdow_b200.benchmarks import MaximumReturn
    x = torch.zeros(3, 1, 4, 10, device="cuda:0")             # all equal: the LP is degenerate, lowest indices are funded
    w = MaximumReturn(max_weight=0.3)(x)
    assert torch.allclose(w, torch.tensor([0.3, 0.3, 0.3, 0.1] + [0.0] * 6, device="cuda:0").expand(3, -1), atol=1e-6)
    with pytest.raises(Exception):
        MaximumReturn(max_weight=0.05)(x)                     # 10 * 0.05 < 1


@pytest.mark.gpu
@pytest.mark.parametrize("N,u", [(10, 1.0), (20, 0.2), (50, 0.05)])
def test_minimum_variance_vs_oracle(N, u):
    from deepdow_b200.benchmarks import MinimumVariance
    rng = np.random.default_rng(N)
    x = torch.tensor(rng.standard_normal((12, 1, 40, N)) * 0.01, dtype=torch.float64, device="cuda:0")
    w = MinimumVariance(max_weight=u, n_assets=N)(x)
    assert w.shape == (12, N) and w.dtype == x.dtype
    A = oracle.covariance_fwd(x[:, 0].cpu().numpy(), 0.5, "diagonal", True)
    ref, *_ = oracle.markowitz_fwd(np.zeros((12, N)), A, np.ones(12), np.zeros(12), u)
    assert np.abs(w.cpu().numpy() - ref).max() < 1e-7
    assert torch.allclose(w.sum(1), torch.ones(12, dtype=w.dtype, device=w.device), atol=1e-9)

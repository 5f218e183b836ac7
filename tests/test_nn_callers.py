"""Callers of the hot path (deepdow_b200/nn.py): the batched RNN / AttentionCollapse against fixtures produced by the
UNMODIFIED reference modules (tests/golden/make_golden_nn.py: deepdow/layers/transform.py, collapse.py, losses.py loaded by
file path), state-dict compatibility with the reference's parameter names, and -- on the GPU -- training steps of BachelierNet
and ThorpNet through the CUDA layers.

The per-sample loops of the reference (transform.py:151-159, collapse.py:46-61) are restated inline below as the second check:
the batched modules must reproduce them on the same parameters.
"""
import os

import numpy as np
import pytest
import torch

from deepdow_b200 import nn as dnn


def _load(golden_dir):
    return np.load(os.path.join(golden_dir, "nn_reference.npz"))


def _state(z, tag):
    pre = f"{tag}_p_"
    return {k[len(pre):]: torch.tensor(z[k]) for k in z.files if k.startswith(pre)}


@pytest.mark.parametrize("tag,kw", [("lstm", dict(cell_type="LSTM", bidirectional=True, n_layers=1)),
                                    ("rnn", dict(cell_type="RNN", bidirectional=False, n_layers=2))])
def test_batched_rnn_matches_reference_fixture(golden_dir, tag, kw):
    z = _load(golden_dir)
    layer = dnn.RNN(2, hidden_size=8, **kw).double()
    layer.load_state_dict(_state(z, tag))            # same parameter names as the reference's RNN.cell
    x = torch.tensor(z[f"{tag}_x"])
    y = layer(x)
    assert y.shape == (3, 8, 5, 4)
    assert torch.allclose(y, torch.tensor(z[f"{tag}_y"]), atol=1e-12)
    # the reference's formulation, sample by sample, on the same cell
    loop = torch.stack([layer.cell(x.permute(0, 2, 3, 1)[i])[0].permute(2, 0, 1) for i in range(len(x))])
    assert torch.allclose(y, loop, atol=1e-12)


def test_batched_attention_matches_reference_fixture(golden_dir):
    z = _load(golden_dir)
    layer = dnn.AttentionCollapse(6).double()
    layer.load_state_dict(_state(z, "att"))
    x = torch.tensor(z["att_x"], requires_grad=True)
    y = layer(x)
    assert y.shape == (3, 6, 4)
    assert torch.allclose(y, torch.tensor(z["att_y"]), atol=1e-12)
    y.sum().backward()
    assert torch.isfinite(x.grad).all()


def test_sharpe_ratio_loss_matches_reference_fixture(golden_dir):
    z = _load(golden_dir)
    loss = dnn.sharpe_ratio_loss(torch.tensor(z["sharpe_w"]), torch.tensor(z["sharpe_y"]))
    assert torch.allclose(loss, torch.tensor(z["sharpe_loss"]), atol=1e-12)


def test_rnn_errors_and_average_collapse():
    with pytest.raises(ValueError):
        dnn.RNN(2, hidden_size=7, bidirectional=True)
    with pytest.raises(ValueError):
        dnn.RNN(2, hidden_size=8, cell_type="GRU")
    x = torch.arange(24.0).reshape(2, 3, 4)
    assert torch.equal(dnn.AverageCollapse(collapse_dim=1)(x), x.mean(1))


def test_network_parameter_names_follow_the_reference():
    # nn.py:129-150 and :521-539: a checkpoint of the reference network loads into these modules
    b = dnn.BachelierNet(2, 6, hidden_size=8)
    names = set(b.state_dict())
    assert {"norm_layer.weight", "norm_layer.bias", "gamma_sqrt", "alpha", "transform_layer.cell.weight_ih_l0",
            "transform_layer.cell.weight_hh_l0_reverse", "time_collapse_layer.affine.weight",
            "time_collapse_layer.context_vector.weight"} <= names
    assert sum(p.numel() for p in b.portfolio_opt_layer.parameters()) == 0
    th = dnn.ThorpNet(5, max_weight=0.5, force_symmetric=False)
    assert set(th.state_dict()) == {"matrix", "exp_returns", "gamma_sqrt", "alpha"}
    assert th.hparams == {"n_assets": 5, "max_weight": 0.5, "force_symmetric": "False"} or th.hparams["n_assets"] == 5
    assert b.hparams["hidden_size"] == 8


# ------------------------------------------------------------------ GPU: the networks on the CUDA layers
@pytest.mark.gpu
def test_bacheliernet_training_steps_config2():
    """BASELINE.json config 2: n_assets=50, lookback=40, horizon=10, batch=256, SharpeRatio loss, Adam (experiments.py:297)."""
    dev = "cuda:0"
    torch.manual_seed(0)
    net = dnn.BachelierNet(1, 50, hidden_size=32, max_weight=1, shrinkage_strategy="diagonal", p=0.5).to(dev)
    opt = torch.optim.Adam(net.parameters(), lr=1e-2)
    X = torch.randn(256, 1, 40, 50, device=dev) * 0.01
    y = torch.randn(256, 1, 10, 50, device=dev) * 0.01
    losses = []
    for _ in range(5):
        w = net(X)
        assert w.shape == (256, 50) and w.dtype == X.dtype and w.device == X.device
        assert torch.allclose(w.sum(1), torch.ones(256, device=dev), atol=1e-4)
        assert (w >= -1e-6).all() and (w <= 1 + 1e-6).all()
        loss = dnn.sharpe_ratio_loss(w, y).mean()
        opt.zero_grad()
        loss.backward()
        for name, p in net.named_parameters():
            assert p.grad is not None and torch.isfinite(p.grad).all(), name
        opt.step()
        losses.append(loss.item())
    assert np.isfinite(losses).all()
    net.eval()
    with torch.no_grad():
        w1, w2 = net(X), net(X)
    assert torch.equal(w1, w2)          # deterministic in eval mode


@pytest.mark.gpu
def test_bacheliernet_layers_match_oracle_inside_the_network():
    """The allocator inside the network sees exactly (exp_rets, covmat, gamma, alpha): recompute its output with the oracle."""
    import oracle
    dev = "cuda:0"
    torch.manual_seed(1)
    net = dnn.BachelierNet(2, 12, hidden_size=8, max_weight=0.3).to(dev).eval()
    X = torch.randn(9, 2, 20, 12, device=dev) * 0.01
    seen = {}
    net.portfolio_opt_layer.register_forward_hook(lambda m, inp, out: seen.update(inp=[i.detach() for i in inp], out=out.detach()))
    net.covariance_layer.register_forward_hook(lambda m, inp, out: seen.update(cov_in=inp[0].detach(), cov_out=out.detach()))
    with torch.no_grad():
        w = net(X)
    rets, cov, g, a = [i.double().cpu().numpy() for i in seen["inp"]]
    w_ref, *_ = oracle.markowitz_fwd(rets, cov, g, a, 0.3)
    assert np.abs(w.double().cpu().numpy() - w_ref).max() < 1e-5
    cov_ref = oracle.covariance_fwd(seen["cov_in"].double().cpu().numpy(), 0.5, "diagonal", False)
    assert np.abs(seen["cov_out"].double().cpu().numpy() - cov_ref).max() < 1e-5 * max(1.0, np.abs(cov_ref).max())


@pytest.mark.gpu
@pytest.mark.parametrize("force_symmetric", [True, False])
def test_thorpnet_training_steps(force_symmetric):
    """nn.py:541-579; tests/test_nn.py:241-278 check shape, dtype, device and the budget -- so do we, plus finite gradients."""
    dev = "cuda:0"
    torch.manual_seed(2)
    net = dnn.ThorpNet(20, max_weight=0.2, force_symmetric=force_symmetric).to(dev)
    with torch.no_grad():
        net.exp_returns.normal_(0, 0.1)
        net.matrix.add_(0.05 * torch.randn(20, 20, device=dev))
    opt = torch.optim.Adam(net.parameters(), lr=1e-2)
    X = torch.randn(64, 1, 5, 20, device=dev)
    y = torch.randn(64, 1, 6, 20, device=dev) * 0.01
    for _ in range(3):
        w = net(X)
        assert w.shape == (64, 20)
        assert torch.allclose(w.sum(1), torch.ones(64, device=dev), atol=1e-4)
        assert (w <= 0.2 + 1e-6).all()
        assert torch.allclose(w, w[:1].expand_as(w), atol=1e-6)      # identical problems, identical answers
        loss = dnn.sharpe_ratio_loss(w, y).mean()
        opt.zero_grad()
        loss.backward()
        assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in net.parameters())
        opt.step()

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


def test_keynesnet_constructor_follows_the_reference():
    # nn.py:246-296: errors, parameter names, the analytical softmax as the default allocator
    with pytest.raises(ValueError):
        dnn.KeynesNet(2, transform_type="GRU")
    with pytest.raises(ValueError):
        dnn.KeynesNet(2, hidden_size=10, n_groups=4)
    net = dnn.KeynesNet(2, hidden_size=8, n_groups=4)
    assert {"temperature", "norm_layer_1.weight", "norm_layer_2.weight", "transform_layer.cell.weight_ih_l0"} <= set(net.state_dict())
    assert net.portfolio_opt_layer.formulation == "analytical" and net.portfolio_opt_layer.temperature is None
    w = net(torch.randn(3, 2, 7, 5))                         # the default allocator is plain torch: runs anywhere
    assert w.shape == (3, 5) and torch.allclose(w.sum(1), torch.ones(3), atol=1e-6)
    wc = dnn.KeynesNet(2, hidden_size=8, transform_type="Conv")(torch.randn(3, 2, 7, 5))
    assert wc.shape == (3, 5)


def test_keynesnet_conv_equals_the_per_asset_loop():
    """nn.py:318-322 convolves every asset separately with one 1-D kernel; the (3, 1) 2-D kernel is the same map."""
    torch.manual_seed(0)
    net = dnn.KeynesNet(2, hidden_size=8, transform_type="Conv").double()
    conv1d = torch.nn.Conv1d(2, 8, kernel_size=3, padding=1).double()
    with torch.no_grad():
        conv1d.weight.copy_(net.transform_layer.weight[..., 0])
        conv1d.bias.copy_(net.transform_layer.bias)
    x = torch.randn(3, 2, 7, 5, dtype=torch.float64)
    loop = torch.stack([conv1d(x[..., i]) for i in range(5)], dim=-1)
    assert torch.allclose(net.transform_layer(x), loop, atol=1e-12)


def test_resample_rejects_other_allocators():
    from deepdow_b200 import Resample
    with pytest.raises(TypeError):                           # allocate.py:323-328
        Resample(torch.nn.Identity())


@pytest.mark.gpu
@pytest.mark.parametrize("kind,max_weight", [("sparsemax", 1.0), ("sparsemax", 0.05), ("softmax", 1.0), ("softmax", 0.05)])
def test_keynesnet_with_cuda_allocators_config3(kind, max_weight):
    """BASELINE.json config 3: KeynesNet with its allocator swapped for the sparsemax / variational softmax, n_assets=100,
    batch=1024 (nn.py:296-298 is where a user swaps it).  The allocator's output inside the network equals the oracle's on
    the logits and temperatures it was given, and a training step reaches every parameter."""
    import oracle
    from deepdow_b200 import SoftmaxAllocator, SparsemaxAllocator
    dev = "cuda:0"
    torch.manual_seed(3)
    alloc = (SparsemaxAllocator(100, temperature=None, max_weight=max_weight) if kind == "sparsemax"
             else SoftmaxAllocator(temperature=None, formulation="variational", n_assets=100, max_weight=max_weight))
    net = dnn.KeynesNet(2, hidden_size=8, n_groups=4, portfolio_opt_layer=alloc).to(dev)
    with torch.no_grad():
        net.temperature.fill_(0.05)                           # sharp enough that the cap and the zero bound bind
    X = torch.randn(1024, 2, 12, 100, device=dev)
    y = torch.randn(1024, 2, 6, 100, device=dev) * 0.01
    seen = {}
    alloc.register_forward_hook(lambda m, inp, out: seen.update(x=inp[0].detach(), t=inp[1].detach(), w=out.detach()))
    w = net(X)
    assert w.shape == (1024, 100) and w.dtype == X.dtype and w.device == X.device
    assert torch.allclose(w.sum(1), torch.ones(1024, device=dev), atol=1e-5)
    assert (w >= -1e-6).all() and (w <= max_weight + 1e-6).all()
    xin, tin = seen["x"].double().cpu().numpy(), seen["t"].double().cpu().numpy()
    ref = (oracle.capped_simplex_fwd if kind == "sparsemax" else oracle.capped_softmax_fwd)(xin, tin, max_weight)
    assert np.abs(seen["w"].double().cpu().numpy() - ref).max() < 1e-5
    if kind == "sparsemax":
        clear = np.abs(ref) > 1e-6                            # supports must match exactly away from ties
        assert ((seen["w"].cpu().numpy() > 0) == (ref > 0))[clear | (ref == 0)].all()
    loss = dnn.sharpe_ratio_loss(w, y).mean()
    loss.backward()
    for name, p in net.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), name
    assert float(net.temperature.grad.abs()) > 0


@pytest.mark.gpu
def test_thorpnet_shared_problem_equals_the_per_sample_path():
    """the n identical problems of a ThorpNet batch solved once (opt-in) give the weights and the
    parameter gradients of the per-sample path (nn.py:563-579)."""
    dev = "cuda:0"
    torch.manual_seed(5)
    net = dnn.ThorpNet(40, max_weight=0.1).to(dev)
    with torch.no_grad():
        net.exp_returns.normal_(0, 0.05)
        net.matrix.add_(0.05 * torch.randn(40, 40, device=dev))
    X = torch.randn(256, 1, 5, 40, device=dev)
    y = torch.randn(256, 1, 6, 40, device=dev) * 0.01
    grads = []
    for shared in (False, True):
        net.shared_problem = shared
        net.zero_grad()
        w = net(X)
        dnn.sharpe_ratio_loss(w, y).mean().backward()
        grads.append((w.detach().clone(), [p.grad.clone() for p in net.parameters()]))
    assert torch.allclose(grads[0][0], grads[1][0], atol=1e-6)
    for ga, gb in zip(grads[0][1], grads[1][1]):
        assert torch.allclose(ga, gb, rtol=2e-4, atol=1e-6 * float(ga.abs().max()) + 1e-9)


@pytest.mark.gpu
def test_bacheliernet_fused_covariance_equals_the_default_path():
    """BachelierNet.fused_covariance (opt-in): same weights bit for bit, same parameter gradients."""
    dev = "cuda:0"
    torch.manual_seed(4)
    net = dnn.BachelierNet(1, 50, hidden_size=16, max_weight=1, p=0.0).to(dev)
    X = torch.randn(64, 1, 40, 50, device=dev) * 0.01
    y = torch.randn(64, 1, 10, 50, device=dev) * 0.01
    out = []
    for fused in (False, True):
        net.fused_covariance = fused
        net.zero_grad()
        w = net(X)
        dnn.sharpe_ratio_loss(w, y).mean().backward()
        out.append((w.detach().clone(), {k: p.grad.clone() for k, p in net.named_parameters()}))
    assert torch.equal(out[0][0], out[1][0])
    for k in out[0][1]:
        assert torch.allclose(out[0][1][k], out[1][1][k], rtol=1e-4, atol=1e-7), k

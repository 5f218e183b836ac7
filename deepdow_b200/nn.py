"""Callers of the allocation hot path: BachelierNet, ThorpNet and KeynesNet on the
CUDA layers.

Only what sits between the data tensor and the allocator is here, and only because the reference's versions serialise the
batch: ``RNN.forward`` (deepdow/layers/transform.py:151-159) and ``AttentionCollapse.forward`` (deepdow/layers/collapse.py:46-61)
loop over the samples in Python, which is what dominates a BachelierNet step once the solver runs on the GPU.  Here the whole
batch goes through the recurrent cell and the attention as ONE call each -- (lookback, n_samples * n_assets, channels) for the
cell, a 4-D einsum-free broadcast for the attention -- which is numerically the same computation: every (sample, asset) pair
is an independent sequence in both formulations.  Module and parameter names follow the reference so that a ``state_dict``
moves between the two implementations unchanged (nn.py:129-150 for BachelierNet, nn.py:521-539 for ThorpNet).
"""
import torch
from torch import nn

from .layers import CovarianceMatrix, NumericalMarkowitz, SoftmaxAllocator, covariance_markowitz


class RNN(nn.Module):
    """``deepdow.layers.RNN`` (transform.py:64-159) with the sample loop folded into the batch dimension of the cell.

    ``(n_samples, n_channels, lookback, n_assets) -> (n_samples, hidden_size, lookback, n_assets)``.
    """

    def __init__(self, n_channels, hidden_size, cell_type="LSTM", bidirectional=True, n_layers=1):
        super().__init__()
        if bidirectional and hidden_size % 2 != 0:
            raise ValueError("Hidden size needs to be divisible by two for bidirectional RNNs.")
        per_direction = hidden_size // (2 if bidirectional else 1)
        cells = {"RNN": nn.RNN, "LSTM": nn.LSTM}
        if cell_type not in cells:
            raise ValueError("Unsupported cell_type {}".format(cell_type))
        self.cell = cells[cell_type](n_channels, per_direction, bidirectional=bidirectional, num_layers=n_layers)

    def forward(self, x):
        n_samples, n_channels, lookback, n_assets = x.shape
        # time first, (sample, asset) pairs as the cell's batch: one launch sequence instead of n_samples
        seq = x.permute(2, 0, 3, 1).reshape(lookback, n_samples * n_assets, n_channels)
        hidden = self.cell(seq)[0]                                   # (lookback, n_samples * n_assets, hidden_size)
        return hidden.reshape(lookback, n_samples, n_assets, -1).permute(1, 3, 0, 2)


class AttentionCollapse(nn.Module):
    """``deepdow.layers.AttentionCollapse`` (collapse.py:8-61) over the whole batch at once.

    ``(n_samples, n_channels, lookback, n_assets) -> (n_samples, n_channels, n_assets)``: a softmax over the lookback of
    ``context_vector(affine(x))`` weighs the time steps; the weighted values are averaged (mean, as in the reference).
    """

    def __init__(self, n_channels):
        super().__init__()
        self.affine = nn.Linear(n_channels, n_channels)
        self.context_vector = nn.Linear(n_channels, 1, bias=False)

    def forward(self, x):
        seq = x.permute(0, 3, 2, 1)                                  # (n_samples, n_assets, lookback, n_channels)
        score = self.context_vector(self.affine(seq))                # (n_samples, n_assets, lookback, 1)
        weight = torch.softmax(score, dim=2)
        return (seq * weight).mean(dim=2).permute(0, 2, 1)           # (n_samples, n_channels, n_assets)


class AverageCollapse(nn.Module):
    """``deepdow.layers.AverageCollapse`` (collapse.py:64-90): mean over one dimension."""

    def __init__(self, collapse_dim=2):
        super().__init__()
        self.collapse_dim = collapse_dim

    def forward(self, x):
        return x.mean(dim=self.collapse_dim)


def _hparams(values):
    return {k: v if isinstance(v, (int, float, str)) else str(v) for k, v in values.items() if k != "self"}


class BachelierNet(nn.Module):
    """``deepdow.nn.BachelierNet`` (nn.py:63-203): instance norm -> LSTM -> attention over time -> channel mean as expected
    returns; sample covariance of the normalised first channel (``sqrt=False``) as the risk matrix; NumericalMarkowitz on top
    with one learnt ``gamma_sqrt`` and ``alpha`` for every sample."""

    def __init__(self, n_input_channels, n_assets, hidden_size=32, max_weight=1, shrinkage_strategy="diagonal", p=0.5):
        super().__init__()
        self._hp = _hparams(dict(n_input_channels=n_input_channels, n_assets=n_assets, hidden_size=hidden_size,
                                 max_weight=max_weight, shrinkage_strategy=shrinkage_strategy, p=p))
        self.norm_layer = nn.InstanceNorm2d(n_input_channels, affine=True)
        self.transform_layer = RNN(n_input_channels, hidden_size=hidden_size)
        self.dropout_layer = nn.Dropout(p=p)
        self.time_collapse_layer = AttentionCollapse(n_channels=hidden_size)
        self.covariance_layer = CovarianceMatrix(sqrt=False, shrinkage_strategy=shrinkage_strategy)
        self.channel_collapse_layer = AverageCollapse(collapse_dim=1)
        self.portfolio_opt_layer = NumericalMarkowitz(n_assets, max_weight=max_weight)
        self.gamma_sqrt = nn.Parameter(torch.ones(1), requires_grad=True)
        self.alpha = nn.Parameter(torch.ones(1), requires_grad=True)
        # Opt-in: form the covariance inside the solver kernel (ddw_covariance_markowitz_fwd) instead of
        # writing (n_samples, n_assets, n_assets) to HBM and reading it back.  Same weights, bit for bit.
        self.fused_covariance = False

    def forward(self, x):
        """x (n_samples, n_channels, lookback, n_assets) -> weights (n_samples, n_assets)."""
        x = self.norm_layer(x)
        h = self.time_collapse_layer(self.dropout_layer(self.transform_layer(x)))
        exp_rets = self.channel_collapse_layer(h)
        ones = torch.ones(len(x), device=x.device, dtype=x.dtype)
        if self.fused_covariance:
            return covariance_markowitz(x[:, 0], exp_rets, ones * self.gamma_sqrt, ones * self.alpha, self.covariance_layer,
                                        self.portfolio_opt_layer)
        covmat = self.covariance_layer(x[:, 0])
        return self.portfolio_opt_layer(exp_rets, covmat, ones * self.gamma_sqrt, ones * self.alpha)

    @property
    def hparams(self):
        return dict(self._hp)


class ThorpNet(nn.Module):
    """``deepdow.nn.ThorpNet`` (nn.py:482-590): every sample solves the same learnt problem -- a learnt matrix (``M M'`` when
    ``force_symmetric``) in the risk slot, a learnt vector of expected returns, ``gamma_sqrt`` and ``alpha``.  The (n, N, N)
    copies the reference builds with ``repeat_interleave`` are built here too: the layer's contract is per-sample inputs, and
    autograd sums the per-sample gradients back into the parameters exactly as in the reference."""

    def __init__(self, n_assets, max_weight=1, force_symmetric=True):
        super().__init__()
        self._hp = _hparams(dict(n_assets=n_assets, max_weight=max_weight, force_symmetric=force_symmetric))
        self.force_symmetric = force_symmetric
        # Opt-in: every sample of a ThorpNet batch is the SAME problem (nn.py:563-575), so it can be
        # solved once and broadcast; autograd of the broadcast sums the n identical upstream gradients, which is exactly what
        # the reference's repeat_interleave backward does.  Off by default: the layer then sees the (n, N, N) copies, as there.
        self.shared_problem = False
        self.matrix = nn.Parameter(torch.eye(n_assets), requires_grad=True)
        self.exp_returns = nn.Parameter(torch.zeros(n_assets), requires_grad=True)
        self.gamma_sqrt = nn.Parameter(torch.ones(1), requires_grad=True)
        self.alpha = nn.Parameter(torch.ones(1), requires_grad=True)
        self.portfolio_opt_layer = NumericalMarkowitz(n_assets, max_weight=max_weight)

    def forward(self, x):
        """x (n_samples, n_channels, lookback, n_assets) -> weights (n_samples, n_assets); only len(x) is used."""
        n = len(x)
        covariance = self.matrix @ self.matrix.t() if self.force_symmetric else self.matrix
        if self.shared_problem:
            one = torch.ones(1, device=x.device, dtype=x.dtype)
            w = self.portfolio_opt_layer(self.exp_returns.unsqueeze(0), covariance.unsqueeze(0), one * self.gamma_sqrt,
                                         one * self.alpha)
            return w.expand(n, -1)
        rets_all = self.exp_returns.unsqueeze(0).expand(n, -1).contiguous()
        cov_all = covariance.unsqueeze(0).expand(n, -1, -1).contiguous()
        ones = torch.ones(n, device=x.device, dtype=x.dtype)
        return self.portfolio_opt_layer(rets_all, cov_all, ones * self.gamma_sqrt, ones * self.alpha)

    @property
    def hparams(self):
        return dict(self._hp)


class KeynesNet(nn.Module):
    """``deepdow.nn.KeynesNet`` (nn.py:205-344): instance norm -> shared LSTM (or 1-D convolution) over time -> group norm ->
    ReLU -> mean over time and channels -> allocator with one learnt temperature.  The reference hard-wires the analytical
    ``SoftmaxAllocator(temperature=None)`` (nn.py:296) and a user swaps ``portfolio_opt_layer`` for the variational softmax or
    the sparsemax (BASELINE config 3); here the layer can be passed to the constructor as well.  Any module with the
    ``forward(x, temperature)`` signature works."""

    def __init__(self, n_input_channels, hidden_size=32, transform_type="RNN", n_groups=4, portfolio_opt_layer=None):
        super().__init__()
        self._hp = _hparams(dict(n_input_channels=n_input_channels, hidden_size=hidden_size, transform_type=transform_type,
                                 n_groups=n_groups))
        self.transform_type = transform_type
        if transform_type == "RNN":
            self.transform_layer = RNN(n_input_channels, hidden_size=hidden_size, bidirectional=False, cell_type="LSTM")
        elif transform_type == "Conv":
            # the reference convolves every asset's (channels, lookback) series with the same 1-D kernel in a Python loop over
            # the assets (nn.py:318-322); a (3, 1) 2-D kernel over (lookback, n_assets) is the same map in one call
            self.transform_layer = nn.Conv2d(n_input_channels, hidden_size, kernel_size=(3, 1), padding=(1, 0))
        else:
            raise ValueError("Unsupported transform_type: {}".format(transform_type))
        if hidden_size % n_groups != 0:
            raise ValueError("The hidden_size needs to be divisible by the n_groups.")
        self.norm_layer_1 = nn.InstanceNorm2d(n_input_channels, affine=True)
        self.temperature = nn.Parameter(torch.ones(1), requires_grad=True)
        self.norm_layer_2 = nn.GroupNorm(n_groups, hidden_size, affine=True)
        self.time_collapse_layer = AverageCollapse(collapse_dim=2)
        self.channel_collapse_layer = AverageCollapse(collapse_dim=1)
        self.portfolio_opt_layer = SoftmaxAllocator(temperature=None) if portfolio_opt_layer is None else portfolio_opt_layer

    def forward(self, x):
        """x (n_samples, n_channels, lookback, n_assets) -> weights (n_samples, n_assets)."""
        x = self.transform_layer(self.norm_layer_1(x))
        x = torch.relu(self.norm_layer_2(x))
        x = self.channel_collapse_layer(self.time_collapse_layer(x))
        temperatures = torch.ones(len(x), device=x.device, dtype=x.dtype) * self.temperature
        return self.portfolio_opt_layer(x, temperatures)

    @property
    def hparams(self):
        return dict(self._hp)


def sharpe_ratio_loss(weights, y, returns_channel=0, rf=0.0, eps=1e-4):
    """Negative Sharpe ratio of the portfolio's simple returns over the horizon -- ``deepdow.losses.SharpeRatio`` with its
    defaults (losses.py:792-799 on top of ``portfolio_returns`` :76-137: log returns in, simple returns out, no rebalancing,
    so after the first step the holdings drift with the assets' cumulative growth and are renormalised).

    weights (n_samples, n_assets), y (n_samples, n_channels, horizon, n_assets) -> (n_samples,)."""
    simple = torch.exp(y[:, returns_channel]) - 1                     # (n_samples, horizon, n_assets)
    drift = (1 + simple).cumprod(dim=1)[:, :-1] * weights.unsqueeze(1)
    held = torch.cat([weights.unsqueeze(1), drift / drift.sum(dim=2, keepdim=True)], dim=1)
    prets = (simple * held).sum(-1)                                   # (n_samples, horizon)
    return -(prets.mean(dim=1) - rf) / (prets.std(dim=1) + eps)

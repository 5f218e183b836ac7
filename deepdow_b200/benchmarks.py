"""The two cvxpylayers-backed benchmarks of the reference on the CUDA path.

``MinimumVariance`` (deepdow/benchmarks.py:166-262) is NumericalMarkowitz's program with ``rets = 0``, ``alpha = 0``,
``gamma = 1`` on ``CovarianceMatrix(sqrt=True)`` of the returns channel; ``MaximumReturn`` (benchmarks.py:76-163) is a linear
program over the capped simplex whose vertex solution the ``ddw_capped_argmax_fwd`` kernel writes directly.  Constructor
arguments, ``__call__`` signature, ``hparams`` and the ``ValueError`` on a wrong asset count follow the reference.  Benchmarks
are inference only (benchmarks.py:13-15); both run under ``torch.no_grad``.
"""
import torch

from . import _C
from .layers import CovarianceMatrix, NumericalMarkowitz


class _Benchmark:
    def _check_assets(self, n_assets):
        if self.n_assets is not None and self.n_assets != n_assets:
            raise ValueError("Incorrect number of assets: {}, expected: {}".format(n_assets, self.n_assets))


class MaximumReturn(_Benchmark):
    """Fully invested, long only, capped portfolio with the highest mean return over the lookback."""

    def __init__(self, max_weight=1, n_assets=None, returns_channel=0):
        self.max_weight = max_weight
        self.n_assets = n_assets
        self.returns_channel = returns_channel

    def __call__(self, x):
        """x (n_samples, n_channels, lookback, n_assets) -> weights (n_samples, n_assets)."""
        n_samples, _, lookback, n_assets = x.shape
        self._check_assets(n_assets)
        with torch.no_grad():
            rets = x[:, self.returns_channel].mean(dim=1).contiguous()
            _C.require_cuda(rets)
            w = torch.empty_like(rets)
            with _C.device_guard(rets):
                rc = _C.lib().ddw_capped_argmax_fwd(_C.ptr(rets), float(self.max_weight), n_samples, n_assets,
                                                    _C.dtype_code(rets), _C.ptr(w), _C.stream(rets.device))
            _C.check(rc, "ddw_capped_argmax_fwd")
        return w

    @property
    def hparams(self):
        return {"max_weight": self.max_weight, "returns_channel": self.returns_channel, "n_assets": self.n_assets}


class MinimumVariance(_Benchmark):
    """Fully invested, long only, capped portfolio of minimum variance under the shrunk sample covariance."""

    def __init__(self, max_weight=1, returns_channel=0, n_assets=None):
        self.n_assets = n_assets
        self.returns_channel = returns_channel
        self.max_weight = max_weight
        self._cov = CovarianceMatrix(sqrt=True)

    def __call__(self, x):
        """x (n_samples, n_channels, lookback, n_assets) -> weights (n_samples, n_assets)."""
        n_samples, _, lookback, n_assets = x.shape
        self._check_assets(n_assets)
        with torch.no_grad():
            covmat_sqrt = self._cov(x[:, self.returns_channel])
            zeros = torch.zeros(n_samples, n_assets, dtype=x.dtype, device=x.device)
            ones = torch.ones(n_samples, dtype=x.dtype, device=x.device)
            layer = NumericalMarkowitz(n_assets, max_weight=self.max_weight)
            # a benchmark is called once per evaluation and the layer is rebuilt per call: check the solver status inside the
            # call (the reference raises SolverError from here too); the lazy mode would never get to report
            layer.check_status = True
            return layer(zeros, covmat_sqrt, ones, zeros[:, 0])

    @property
    def hparams(self):
        return {"max_weight": self.max_weight, "returns_channel": self.returns_channel, "n_assets": self.n_assets}

"""Host-buffer path of NumericalMarkowitz: CPU tensors in, CPU tensors out, solved on the GPU.

The reference layer (deepdow/layers/allocate.py:240-273) is called with CPU tensors and returns CPU
tensors.  ``MarkowitzHostPipeline`` serves such a caller through ``ddw_markowitz_host_step``
(include/ddw.h): the batch is streamed through the GPU in chunks on three CUDA streams, so the
host->device copies, the solver kernels and the device->host copies overlap.  There is still no CPU
compute: the solver runs on the GPU only.
"""
import collections
import ctypes

import torch

from . import _C
from .layers import SolverError


HostStepResult = collections.namedtuple("HostStepResult",
                                        ["w", "d_rets", "d_covmat_sqrt", "d_gamma_sqrt", "d_alpha", "status"])
# the same with the gradient of covmat_sqrt in factored form: factors (B, 2, N) = (A d, A w), see expand_factored()
HostStepFactored = collections.namedtuple("HostStepFactored",
                                          ["w", "d_rets", "factors", "d_gamma_sqrt", "d_alpha", "status"])


def expand_factored(result, gamma_sqrt):
    """d_covmat_sqrt (B, N, N) from a ``HostStepFactored``: ``-2 gamma_ * (p w' + q d')`` with ``(p, q) = factors`` and
    ``gamma_`` the reference's tiling of ``gamma_sqrt`` (allocate.py:268-270).  Host-side helper for callers (and tests) that
    want the dense gradient after all; the point of the factored form is that most do not."""
    B, N = result.w.shape
    p, q = result.factors[:, 0], result.factors[:, 1]
    gamma_ = gamma_sqrt.reshape(-1).to(result.w.dtype).repeat((1, N * N)).view(B, N, N)
    return -2 * gamma_ * (p[:, :, None] * result.w[:, None, :] + q[:, :, None] * result.d_rets[:, None, :])


class MarkowitzHostPipeline:
    """Forward (+ backward) of NumericalMarkowitz on host-resident batches.

    Parameters
    ----------
    n_assets : int
    max_weight : float
    chunk : int
        Problems per pipeline stage.  Smaller chunks overlap better, larger ones launch less.
    device : torch.device or int
        GPU that does the work.
    """

    def __init__(self, n_assets, max_weight=1, chunk=128, device=None):
        self.n_assets = int(n_assets)
        self.max_weight = float(max_weight)
        self.chunk = int(chunk)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._handle = None
        self._scratch = None

    # ---------------------------------------------------------------------------------------
    def _ensure(self, B, dt):
        lib = _C.lib()
        if self._handle is None:
            h = ctypes.c_void_p()
            with torch.cuda.device(self.device):
                _C.check(lib.ddw_host_pipeline_create(ctypes.byref(h)), "ddw_host_pipeline_create")
            self._handle = h
        need = lib.ddw_markowitz_host_scratch_bytes(B, self.n_assets, dt, self.chunk)
        if self._scratch is None or self._scratch.numel() < need:
            self._scratch = torch.empty(need, dtype=torch.uint8, device=self.device)
        return lib

    @staticmethod
    def _host(t, dtype, shape, name):
        if t.is_cuda:
            raise ValueError(f"{name}: the host pipeline takes CPU tensors (use the nn.Module for CUDA tensors)")
        t = t.to(dtype).reshape(shape).contiguous()
        return t

    def step(self, rets, covmat_sqrt, gamma_sqrt, alpha, grad_w=None, out=None, sync=True, factored=False):
        """Solve the batch; with ``grad_w`` also return the gradients of ``(w * grad_w).sum()``.

        All arguments are CPU tensors shaped as in the reference's ``forward`` (allocate.py:240-266);
        pinned memory lets the copies run asynchronously.  Returns a ``HostStepResult`` of pinned CPU tensors
        (gradient fields are None without ``grad_w``); pass an earlier result as ``out`` to reuse its buffers.  With ``sync=False`` the call only enqueues; read the results
        after ``torch.cuda.current_stream().synchronize()`` and call ``check()``.  ``factored=True`` (with ``grad_w``) returns a
        ``HostStepFactored``: the rank-2 gradient of covmat_sqrt as its two factor vectors per portfolio (2 N numbers instead
        of N^2 come back over PCIe), see ``expand_factored``.
        """
        B, N = rets.shape
        if N != self.n_assets:
            raise ValueError(f"expected n_assets={self.n_assets}, got {N}")
        if self.n_assets * self.max_weight < 1.0 - 1e-6:
            raise SolverError(f"infeasible: n_assets * max_weight = {self.n_assets * self.max_weight:g} < 1")
        dtype = rets.dtype
        dt = _C.dtype_code(rets)
        lib = self._ensure(B, dt)
        rets_h = self._host(rets, dtype, (B, N), "rets")
        cov_h = self._host(covmat_sqrt, dtype, (B, N, N), "covmat_sqrt")
        gam_h = self._host(gamma_sqrt, dtype, (B,), "gamma_sqrt")
        alp_h = self._host(alpha, dtype, (B,), "alpha")
        g_h = None if grad_w is None else self._host(grad_w, dtype, (B, N), "grad_w")
        factored = bool(factored) and g_h is not None
        if out is None:
            pin = lambda *s, d=dtype: torch.empty(s, dtype=d).pin_memory()
            grads = (pin(B, N), pin(B, 2, N) if factored else pin(B, N, N), pin(B), pin(B)) if g_h is not None else (None,) * 4
            out = (HostStepFactored if factored else HostStepResult)(pin(B, N), *grads, pin(B, d=torch.int32))
        w, status, grads = out.w, out.status, tuple(out[1:5])
        self.last_status = status
        p = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
        d_rets, d_cov, d_gam, d_alp = grads
        fn = lib.ddw_markowitz_host_step_factored if factored else lib.ddw_markowitz_host_step
        with torch.cuda.device(self.device):
            rc = fn(self._handle, p(rets_h), p(cov_h), p(gam_h), p(alp_h), p(g_h),
                    self.max_weight, B, N, dt, self.chunk, p(w), p(status),
                    p(d_rets), p(d_cov), p(d_gam), p(d_alp),
                    p(self._scratch), self._scratch.numel(), _C.stream(self.device))
        _C.check(rc, "ddw_markowitz_host_step")
        self._keep = (rets_h, cov_h, gam_h, alp_h, g_h)   # the copies read these until the stream has drained
        if sync:
            torch.cuda.current_stream(self.device).synchronize()
            self.check()
        return out

    def check(self):
        """Raise SolverError if a problem of the latest (finished) step could not be solved."""
        bad = int(((self.last_status & (_C.STATUS_INEXACT | _C.STATUS_INFEASIBLE)) != 0).sum())
        if bad:
            raise SolverError(f"{bad} portfolio problems could not be solved")

    @property
    def graph_launches(self):
        """Steps served by replaying the cached CUDA graph (same buffers requested again)."""
        return 0 if self._handle is None else int(_C.lib().ddw_host_pipeline_graph_launches(self._handle))

    def close(self):
        if self._handle is not None:
            _C.lib().ddw_host_pipeline_destroy(self._handle)
            self._handle = None
        self._scratch = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

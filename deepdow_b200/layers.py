"""Host-side mirror of the reference layers: thin ``torch.autograd.Function``s over the C ABI.

Argument meaning, error behaviour and output dtype/device follow the reference:
  NumericalMarkowitz   deepdow/layers/allocate.py:198-273
  SoftmaxAllocator     deepdow/layers/allocate.py:414-514
  SparsemaxAllocator   deepdow/layers/allocate.py:517-595
  CovarianceMatrix     deepdow/layers/misc.py:33-201
  NumericalRiskBudgeting  deepdow/layers/allocate.py:631-691
  Resample             deepdow/layers/allocate.py:276-411 (the caller that chains CovarianceMatrix(sqrt=True) into
                       NumericalMarkowitz; only its NumericalMarkowitz arm, the other allocators are out of scope)
"""
import warnings

import torch
import torch.nn as nn

from . import _C

try:  # deepdow's training loop catches diffcp.SolverError (deepdow/experiments.py:6,384)
    from diffcp import SolverError as _BaseSolverError
except Exception:  # diffcp is not a dependency of this package
    _BaseSolverError = Exception


class SolverError(_BaseSolverError):
    """A problem of the batch could not be solved (infeasible box, or no KKT point reached)."""


def _workspace(nbytes, device):
    return torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=device)


# ------------------------------------------------------------------------------------------------
# NumericalMarkowitz
# ------------------------------------------------------------------------------------------------
class _StatusWatch:
    """Deferred read of the solver's failure counter (workspace header, include/ddw.h).

    Reading it inside ``forward`` is a host<->device sync that stalls the launch pipeline (measured:
    0.26 ms of a 0.93 ms fwd+bwd step at B=4096, N=100).  In "lazy" mode the counter is copied
    asynchronously to pinned host memory and inspected by the next call that finds the copy finished,
    so a failure raises ``SolverError`` at most one call late instead of blocking every call.
    """
    SLOTS = 8

    def __init__(self):
        self.buf = None
        self.views = None
        self.events = None
        self.last_header = None
        self.pending = []   # (slot, event, description)
        self.next = 0

    def push(self, counter, what):
        if self.buf is None:
            self.buf = torch.zeros(self.SLOTS, dtype=torch.int32).pin_memory()
            self.views = [self.buf[i:i + 1] for i in range(self.SLOTS)]
            self.events = [torch.cuda.Event() for _ in range(self.SLOTS)]
        if len(self.pending) >= self.SLOTS:
            self.poll(block=True)
        slot = self.next
        self.next = (self.next + 1) % self.SLOTS
        self.views[slot].copy_(counter, non_blocking=True)
        ev = self.events[slot]      # free again: its previous use has been polled (pending < SLOTS)
        ev.record()
        self.pending.append((slot, ev, what))

    def poll(self, block=False):
        while self.pending:
            slot, ev, what = self.pending[0]
            if not ev.query():
                if not block:
                    return
                ev.synchronize()
            self.pending.pop(0)
            bad = int(self.buf[slot])
            if bad:
                self.pending.clear()
                raise SolverError(f"{bad} portfolio problems could not be solved ({what})")


class _MarkowitzFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rets, covmat_sqrt, gamma_sqrt, alpha, max_weight, check_status, watch):
        _C.require_cuda(rets, covmat_sqrt, gamma_sqrt, alpha)
        lib = _C.lib()
        B, N = rets.shape
        dt = _C.dtype_code(rets)
        rets_c, cov_c = _C.aligned(rets), _C.aligned(covmat_sqrt)
        gam_c, alp_c = _C.aligned(gamma_sqrt), _C.aligned(alpha)
        w = torch.empty_like(rets_c)
        state = torch.empty((B, N), dtype=torch.int8, device=rets.device)
        status = torch.empty((B,), dtype=torch.int32, device=rets.device)
        ws = _workspace(_C.markowitz_workspace_bytes(B, N, dt), rets.device)
        # when a backward will follow, the forward keeps the factor of every converged reduced KKT system (dense kernels) in
        # `saved`, and the backward runs substitutions only (include/ddw.h: ddw_markowitz_fwd_saved / _bwd_saved)
        saved = None
        if any(ctx.needs_input_grad[:4]):
            saved = _workspace(lib.ddw_markowitz_saved_bytes(B, N, dt), rets.device)
        with _C.device_guard(rets):
            if saved is None:
                rc = lib.ddw_markowitz_fwd(_C.ptr(rets_c), _C.ptr(cov_c), _C.ptr(gam_c), _C.ptr(alp_c),
                                           float(max_weight), B, N, dt, _C.ptr(w), _C.ptr(state), _C.ptr(status),
                                           _C.ptr(ws), ws.numel(), _C.stream(rets.device))
            else:
                rc = lib.ddw_markowitz_fwd_saved(_C.ptr(rets_c), _C.ptr(cov_c), _C.ptr(gam_c), _C.ptr(alp_c),
                                                 float(max_weight), B, N, dt, _C.ptr(w), _C.ptr(state), _C.ptr(status),
                                                 _C.ptr(ws), ws.numel(), _C.ptr(saved), saved.numel(), _C.stream(rets.device))
        _C.check(rc, "ddw_markowitz_fwd")
        what = f"batch of {B}, n_assets={N}, max_weight={max_weight}"
        counter = ws[8:12].view(torch.int32)          # counter [2] of the workspace header (include/ddw.h)
        watch.last_header = ws[:16].view(torch.int32)  # [fallback, dense hand-over, unsolved, -] of this call (device)
        if check_status is True:
            bad = int(counter)
            if bad:
                raise SolverError(f"{bad} portfolio problems could not be solved ({what})")
        elif check_status == "lazy":
            with _C.device_guard(rets):              # the event is recorded on this tensor's device
                watch.push(counter, what)
        ctx.save_for_backward(w, state, cov_c, gam_c, alp_c, saved)
        ctx.max_weight = float(max_weight)
        ctx.mark_non_differentiable(status)
        return w, status

    @staticmethod
    def backward(ctx, grad_w, _grad_status):
        w, state, cov_c, gam_c, alp_c, saved = ctx.saved_tensors
        lib = _C.lib()
        B, N = w.shape
        dt = _C.dtype_code(w)
        g = _C.aligned(grad_w)
        d_rets = torch.empty_like(w)
        d_cov = torch.empty_like(cov_c)
        d_alpha = torch.empty_like(alp_c)
        need_gamma = ctx.needs_input_grad[2]
        d_gamma = torch.empty_like(gam_c) if need_gamma else None
        ws = _workspace(_C.markowitz_workspace_bytes(B, N, dt), w.device)
        with _C.device_guard(w):
            if saved is None:
                rc = lib.ddw_markowitz_bwd(_C.ptr(g), _C.ptr(w), _C.ptr(state), _C.ptr(cov_c), _C.ptr(gam_c), _C.ptr(alp_c),
                                           ctx.max_weight, B, N, dt, _C.ptr(d_rets), _C.ptr(d_cov), _C.ptr(d_gamma),
                                           _C.ptr(d_alpha), _C.ptr(ws), ws.numel(), _C.stream(w.device))
            else:
                rc = lib.ddw_markowitz_bwd_saved(_C.ptr(g), _C.ptr(w), _C.ptr(state), _C.ptr(cov_c), _C.ptr(gam_c),
                                                 _C.ptr(alp_c), ctx.max_weight, B, N, dt, _C.ptr(d_rets), _C.ptr(d_cov),
                                                 _C.ptr(d_gamma), _C.ptr(d_alpha), _C.ptr(ws), ws.numel(), _C.ptr(saved),
                                                 saved.numel(), _C.stream(w.device))
        _C.check(rc, "ddw_markowitz_bwd")
        return d_rets, d_cov, d_gamma, d_alpha, None, None, None


class _MarkowitzHostFn(torch.autograd.Function):
    """CPU tensors in, CPU tensors out, solved on the GPU through the host-buffer pipeline (opt-in, see
    ``NumericalMarkowitz.host_device``).  The backward re-streams the batch (forward + backward in one pass of
    ``ddw_markowitz_host_step``): the step is PCIe bound, so recomputing the forward costs no extra transfer of w."""

    @staticmethod
    def forward(ctx, rets, covmat_sqrt, gamma_sqrt, alpha, pipe):
        out = pipe.step(rets, covmat_sqrt, gamma_sqrt, alpha)
        ctx.save_for_backward(rets, covmat_sqrt, gamma_sqrt, alpha)
        ctx.pipe = pipe
        ctx.mark_non_differentiable(out.status)
        return out.w, out.status

    @staticmethod
    def backward(ctx, grad_w, _grad_status):
        rets, covmat_sqrt, gamma_sqrt, alpha = ctx.saved_tensors
        out = ctx.pipe.step(rets, covmat_sqrt, gamma_sqrt, alpha, grad_w.contiguous())
        return out.d_rets, out.d_covmat_sqrt, out.d_gamma_sqrt, out.d_alpha, None


class NumericalMarkowitz(nn.Module):
    """Convex optimization layer stylized into portfolio optimization problem.

    Drop-in for ``deepdow.layers.NumericalMarkowitz`` (allocate.py:198-273).  Solves, per sample,

        maximize  rets @ w - ||(gamma_sqrt_ * covmat_sqrt) @ w||^2 - |alpha| * ||w||^2
        s.t.      sum(w) == 1, 0 <= w <= max_weight

    exactly (primal-dual active set on the GPU) instead of through cvxpylayers/SCS.

    Parameters
    ----------
    n_assets : int
        Number of assets.
    max_weight : float
        Upper bound of every weight.

    Attributes
    ----------
    last_status : torch.Tensor or None
        int32 status bits per problem of the latest forward (see ``include/ddw.h``).
    """

    def __init__(self, n_assets, max_weight=1):
        super().__init__()
        self.n_assets = n_assets
        self.max_weight = max_weight
        # "lazy": a failed solve raises SolverError from the next call that finds the status copy finished
        # (no host sync per forward); True: check inside every forward (blocks); False: never check
        self.check_status = "lazy"
        self.last_status = None
        self._watch = _StatusWatch()
        # Opt-in for callers whose tensors live on the host (the reference takes CPU tensors): set to a CUDA device
        # ("cuda:0") and CPU inputs are streamed through that GPU by the host-buffer pipeline (deepdow_b200.host),
        # CPU tensors come back.  None (default): CPU tensors are rejected -- there is no CPU compute path.
        self.host_device = None
        self._host_pipe = None

    def forward(self, rets, covmat_sqrt, gamma_sqrt, alpha):
        """Same arguments as the reference (allocate.py:240-273).

        rets (n_samples, n_assets); covmat_sqrt (n_samples, n_assets, n_assets); gamma_sqrt (n_samples,);
        alpha (n_samples,).  Returns weights (n_samples, n_assets), same dtype and device.
        """
        n_samples, n_assets = rets.shape
        if n_assets != self.n_assets:
            raise ValueError(f"expected n_assets={self.n_assets}, got {n_assets}")
        if covmat_sqrt.shape != (n_samples, n_assets, n_assets):
            raise ValueError("covmat_sqrt needs shape (n_samples, n_assets, n_assets)")
        # the reference's repeat/view needs exactly n_samples gamma values (allocate.py:268-270)
        if gamma_sqrt.numel() != n_samples or alpha.numel() != n_samples:
            raise ValueError("gamma_sqrt and alpha need shape (n_samples,)")
        if self.check_status:
            self._watch.poll()                     # surfaces a failure of an earlier call, never blocks
            if n_assets * float(self.max_weight) < 1.0 - 1e-6:
                # the box leaves no fully invested portfolio: the reference's solver raises here too
                raise SolverError(f"infeasible: n_assets * max_weight = {n_assets * float(self.max_weight):g} < 1")
        dtype = rets.dtype
        if self.host_device is not None and not rets.is_cuda:
            from .host import MarkowitzHostPipeline        # late import: host.py imports this module
            if self._host_pipe is None or self._host_pipe.device != torch.device(self.host_device):
                self._host_pipe = MarkowitzHostPipeline(self.n_assets, self.max_weight, device=self.host_device)
            w, status = _MarkowitzHostFn.apply(rets, covmat_sqrt.to(dtype), gamma_sqrt.reshape(-1).to(dtype),
                                               alpha.reshape(-1).to(dtype), self._host_pipe)
            self.last_status = status
            return w
        w, status = _MarkowitzFn.apply(rets, covmat_sqrt.to(dtype), gamma_sqrt.reshape(-1).to(dtype),
                                       alpha.reshape(-1).to(dtype), self.max_weight, self.check_status, self._watch)
        self.last_status = status
        return w

    def solver_counters(self):
        """(problems sent to the interior-point fallback, problems handed from the sparse-support kernel to the
        dense kernel, problems left unsolved) of the latest forward; synchronises with the device."""
        h = self._watch.last_header
        return None if h is None else tuple(int(v) for v in h[:3].tolist())

    def synchronize_status(self):
        """Block until every outstanding solve has reported; raise SolverError if any failed."""
        self._watch.poll(block=True)

    def train(self, mode=True):
        """``model.train()`` / ``model.eval()`` are the natural synchronisation points of a training loop (deepdow's
        ``Run.launch`` validates after every epoch, callbacks.py:807-836): in the lazy mode the outstanding status reads are
        flushed here, so a solve that failed in the LAST batch of an epoch raises ``SolverError`` at the switch instead of
        going unreported."""
        if self.check_status == "lazy" and self._watch is not None and self._watch.pending:
            self._watch.poll(block=True)
        return super().train(mode)

    def __del__(self):
        watch = self.__dict__.get("_watch")
        if watch is not None and watch.pending:
            try:
                watch.poll(block=True)
            except SolverError as exc:      # cannot raise from a finaliser: make the loss of the report visible at least
                warnings.warn(f"NumericalMarkowitz discarded with an unreported solver failure: {exc}")
            except Exception:               # interpreter shutdown: CUDA may already be gone
                pass

    def extra_repr(self):
        return f"n_assets={self.n_assets}, max_weight={self.max_weight}"

    def __getstate__(self):
        state = self.__dict__.copy()
        state["last_status"] = None
        state["_watch"] = None            # pinned buffers and CUDA events are not picklable
        state["_host_pipe"] = None        # streams and device scratch are rebuilt on demand
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)
        self._watch = _StatusWatch()


# ------------------------------------------------------------------------------------------------
# NumericalRiskBudgeting
# ------------------------------------------------------------------------------------------------
class _RiskBudgetingFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, covmat_sqrt, b, max_weight):
        _C.require_cuda(covmat_sqrt, b)
        lib = _C.lib()
        B, N = b.shape
        dt = _C.dtype_code(b)
        cov_c, b_c = _C.aligned(covmat_sqrt), _C.aligned(b)
        w = torch.empty_like(b_c)
        state = torch.empty((B, N), dtype=torch.int8, device=b.device)
        status = torch.empty((B,), dtype=torch.int32, device=b.device)
        ws = _workspace(lib.ddw_risk_budgeting_workspace_bytes(B, N, dt), b.device)
        with _C.device_guard(b):
            rc = lib.ddw_risk_budgeting_fwd(_C.ptr(cov_c), _C.ptr(b_c), float(max_weight), B, N, dt, _C.ptr(w), _C.ptr(state),
                                            _C.ptr(status), _C.ptr(ws), ws.numel(), _C.stream(b.device))
        _C.check(rc, "ddw_risk_budgeting_fwd")
        bad = int(ws[:4].view(torch.int32))          # the reference raises from inside the call; so do we (one sync)
        if bad:
            raise SolverError(f"{bad} risk budgeting problems could not be solved (batch of {B}, n_assets={N}, "
                              f"max_weight={max_weight})")
        ctx.save_for_backward(w, state, cov_c, b_c)
        ctx.max_weight = float(max_weight)
        ctx.mark_non_differentiable(status)
        return w, status

    @staticmethod
    def backward(ctx, grad_w, _grad_status):
        w, state, cov_c, b_c = ctx.saved_tensors
        lib = _C.lib()
        B, N = w.shape
        dt = _C.dtype_code(w)
        g = _C.aligned(grad_w)
        d_cov = torch.empty_like(cov_c)
        d_b = torch.empty_like(b_c)
        ws = _workspace(lib.ddw_risk_budgeting_workspace_bytes(B, N, dt), w.device)
        with _C.device_guard(w):
            rc = lib.ddw_risk_budgeting_bwd(_C.ptr(g), _C.ptr(w), _C.ptr(state), _C.ptr(cov_c), _C.ptr(b_c), ctx.max_weight, B, N,
                                            dt, _C.ptr(d_cov), _C.ptr(d_b), _C.ptr(ws), ws.numel(), _C.stream(w.device))
        _C.check(rc, "ddw_risk_budgeting_bwd")
        return d_cov, d_b, None


class NumericalRiskBudgeting(nn.Module):
    """Convex optimization layer stylized into portfolio optimization problem.

    Drop-in for ``deepdow.layers.NumericalRiskBudgeting`` (allocate.py:631-691).  Solves, per sample,

        minimize  0.5 * ||covmat_sqrt @ w||^2 - b @ log(w)
        s.t.      sum(w) == 1, 0 <= w <= max_weight

    (at the optimum with a loose cap the risk contribution of every asset is proportional to its budget ``b``) by a
    feasible-point Newton method on the GPU instead of through cvxpylayers.

    Parameters
    ----------
    n_assets : int
        Number of assets.
    max_weight : float
        Upper bound of every weight.
    """

    def __init__(self, n_assets, max_weight=1):
        super().__init__()
        self.n_assets = n_assets
        self.max_weight = max_weight
        self.last_status = None

    def forward(self, covmat_sqrt, b):
        """covmat_sqrt (n_samples, n_assets, n_assets); b (n_samples, n_assets) risk budgets.  Returns weights
        (n_samples, n_assets), same dtype and device."""
        n_samples, n_assets = b.shape
        if n_assets != self.n_assets:
            raise ValueError(f"expected n_assets={self.n_assets}, got {n_assets}")
        if covmat_sqrt.shape != (n_samples, n_assets, n_assets):
            raise ValueError("covmat_sqrt needs shape (n_samples, n_assets, n_assets)")
        if n_assets * float(self.max_weight) < 1.0 - 1e-6:
            raise SolverError(f"infeasible: n_assets * max_weight = {n_assets * float(self.max_weight):g} < 1")
        w, status = _RiskBudgetingFn.apply(covmat_sqrt.to(b.dtype), b, self.max_weight)
        self.last_status = status
        return w

    def extra_repr(self):
        return f"n_assets={self.n_assets}, max_weight={self.max_weight}"

    def __getstate__(self):
        state = self.__dict__.copy()
        state["last_status"] = None
        return state


# ------------------------------------------------------------------------------------------------
# Sparsemax / variational softmax
# ------------------------------------------------------------------------------------------------
class _RowFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, temperature, temperature_scalar, max_weight, kind):
        _C.require_cuda(x, temperature)
        lib = _C.lib()
        B, N = x.shape
        dt = _C.dtype_code(x)
        x_c = x.contiguous()
        t_c = None if temperature is None else temperature.to(x.dtype).contiguous()
        w = torch.empty_like(x_c)
        fwd = lib.ddw_capped_simplex_fwd if kind == "sparsemax" else lib.ddw_capped_softmax_fwd
        with _C.device_guard(x):
            rc = fwd(_C.ptr(x_c), _C.ptr(t_c), float(temperature_scalar), float(max_weight), B, N, dt, _C.ptr(w),
                     _C.stream(x.device))
        _C.check(rc, f"ddw_capped_{kind}_fwd")
        ctx.save_for_backward(w, x_c, t_c)
        ctx.args = (float(temperature_scalar), float(max_weight), kind)
        return w

    @staticmethod
    def backward(ctx, grad_w):
        w, x_c, t_c = ctx.saved_tensors
        tscalar, max_weight, kind = ctx.args
        lib = _C.lib()
        B, N = w.shape
        dt = _C.dtype_code(w)
        g = grad_w.contiguous()
        d_x = torch.empty_like(w)
        d_t = torch.empty_like(t_c) if (t_c is not None and ctx.needs_input_grad[1]) else None
        bwd = lib.ddw_capped_simplex_bwd if kind == "sparsemax" else lib.ddw_capped_softmax_bwd
        with _C.device_guard(w):
            rc = bwd(_C.ptr(g), _C.ptr(w), _C.ptr(x_c), _C.ptr(t_c), tscalar, max_weight, B, N, dt, _C.ptr(d_x),
                     _C.ptr(d_t), _C.stream(w.device))
        _C.check(rc, f"ddw_capped_{kind}_bwd")
        return d_x, d_t, None, None, None


def _row_forward(module, x, temperature, kind):
    # same check as allocate.py:498-499 / :583-584
    if not ((temperature is None) ^ (module.temperature is None)):
        raise ValueError("Not clear which temperature to use")
    if temperature is not None:
        if temperature.numel() != x.shape[0]:
            raise ValueError("temperature needs shape (n_samples,)")
        return _RowFn.apply(x, temperature.reshape(-1), 1.0, module.max_weight, kind)
    return _RowFn.apply(x, None, float(module.temperature), module.max_weight, kind)


class SoftmaxAllocator(nn.Module):
    """Portfolio creation by computing a softmax over the asset dimension with temperature.

    Drop-in for ``deepdow.layers.SoftmaxAllocator`` (allocate.py:414-514).  ``formulation='analytical'``
    is ``torch.nn.Softmax`` as in the reference; ``'variational'`` (entropy-regularised program with a
    weight cap, allocate.py:470-475) runs the capped-softmax kernel.
    """

    def __init__(self, temperature=1, formulation="analytical", n_assets=None, max_weight=1):
        super().__init__()
        self.temperature = temperature
        if formulation not in {"analytical", "variational"}:
            raise ValueError("Unrecognized formulation {}".format(formulation))
        if formulation == "variational" and n_assets is None:
            raise ValueError("One needs to provide n_assets for the variational formulation.")
        if formulation == "analytical" and max_weight != 1:
            raise ValueError("Cannot constraint weights via max_weight for analytical formulation")
        if formulation == "variational" and n_assets * max_weight < 1:
            raise ValueError("One cannot create fully invested portfolio with the given max_weight")
        self.formulation = formulation
        self.n_assets = n_assets
        self.max_weight = max_weight
        if formulation == "analytical":
            self.layer = torch.nn.Softmax(dim=1)

    def forward(self, x, temperature=None):
        """x (n_samples, n_assets); temperature None or (n_samples,).  Returns weights (n_samples, n_assets)."""
        if self.formulation == "analytical":
            n_samples, _ = x.shape
            if not ((temperature is None) ^ (self.temperature is None)):
                raise ValueError("Not clear which temperature to use")
            if temperature is not None:
                temperature_ = temperature
            else:
                temperature_ = float(self.temperature) * torch.ones(n_samples, dtype=x.dtype, device=x.device)
            return self.layer(x / temperature_[..., None])
        if x.shape[1] != self.n_assets:
            raise ValueError(f"expected n_assets={self.n_assets}, got {x.shape[1]}")
        return _row_forward(self, x, temperature, "softmax")


class SparsemaxAllocator(nn.Module):
    """Portfolio creation by computing a sparsemax over the asset dimension with temperature.

    Drop-in for ``deepdow.layers.SparsemaxAllocator`` (allocate.py:517-595): Euclidean projection of
    ``x / temperature`` onto ``{sum(w) = 1, 0 <= w <= max_weight}``.
    """

    def __init__(self, n_assets, temperature=1, max_weight=1):
        super().__init__()
        if n_assets * max_weight < 1:
            raise ValueError("One cannot create fully invested portfolio with the given max_weight")
        self.n_assets = n_assets
        self.temperature = temperature
        self.max_weight = max_weight

    def forward(self, x, temperature=None):
        """x (n_samples, n_assets); temperature None or (n_samples,).  Returns weights (n_samples, n_assets)."""
        if x.shape[1] != self.n_assets:
            raise ValueError(f"expected n_assets={self.n_assets}, got {x.shape[1]}")
        return _row_forward(self, x, temperature, "sparsemax")


# ------------------------------------------------------------------------------------------------
# CovarianceMatrix
# ------------------------------------------------------------------------------------------------
class _CovarianceFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, coef, strategy, do_sqrt):
        _C.require_cuda(x, coef)
        lib = _C.lib()
        B, L, N = x.shape
        dt = _C.dtype_code(x)
        x_c = _C.aligned(x)
        coef_c = None if coef is None else _C.aligned(coef.to(x.dtype))
        out = torch.empty((B, N, N), dtype=x.dtype, device=x.device)
        need_grad = x.requires_grad or (coef is not None and coef.requires_grad)
        eigvec = eigval = None
        if do_sqrt and need_grad:
            eigvec = torch.empty((B, N, N), dtype=x.dtype, device=x.device)
            eigval = torch.empty((B, N), dtype=x.dtype, device=x.device)
        ws = _workspace(lib.ddw_covariance_workspace_bytes(B, L, N, dt, int(do_sqrt)), x.device)
        with _C.device_guard(x):
            rc = lib.ddw_covariance_fwd(_C.ptr(x_c), _C.ptr(coef_c), strategy, int(do_sqrt), B, L, N, dt, _C.ptr(out),
                                        _C.ptr(eigvec), _C.ptr(eigval), None, _C.ptr(ws), ws.numel(), _C.stream(x.device))
        _C.check(rc, "ddw_covariance_fwd")
        ctx.save_for_backward(x_c, coef_c, eigvec, eigval)
        ctx.args = (strategy, int(do_sqrt))
        return out

    @staticmethod
    def backward(ctx, grad_out):
        x_c, coef_c, eigvec, eigval = ctx.saved_tensors
        strategy, do_sqrt = ctx.args
        lib = _C.lib()
        B, L, N = x_c.shape
        dt = _C.dtype_code(x_c)
        g = _C.aligned(grad_out)
        d_x = torch.empty_like(x_c)
        d_coef = torch.empty_like(coef_c) if (coef_c is not None and ctx.needs_input_grad[1]) else None
        ws = _workspace(lib.ddw_covariance_workspace_bytes(B, L, N, dt, do_sqrt), x_c.device)
        with _C.device_guard(x_c):
            rc = lib.ddw_covariance_bwd(_C.ptr(g), _C.ptr(x_c), _C.ptr(coef_c), strategy, do_sqrt, _C.ptr(eigvec),
                                        _C.ptr(eigval), B, L, N, dt, _C.ptr(d_x), _C.ptr(d_coef), _C.ptr(ws),
                                        ws.numel(), _C.stream(x_c.device))
        _C.check(rc, "ddw_covariance_bwd")
        return d_x, d_coef, None, None


class _PsdSqrtFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, m):
        _C.require_cuda(m)
        lib = _C.lib()
        B, N, _ = m.shape
        dt = _C.dtype_code(m)
        m_c = _C.aligned(m)
        out = torch.empty_like(m_c)
        eigvec = eigval = None
        if m.requires_grad:
            eigvec = torch.empty_like(m_c)
            eigval = torch.empty((B, N), dtype=m.dtype, device=m.device)
        ws = _workspace(lib.ddw_covariance_workspace_bytes(B, 2, N, dt, 1), m.device)
        with _C.device_guard(m):
            rc = lib.ddw_psd_sqrt_fwd(_C.ptr(m_c), B, N, dt, _C.ptr(out), _C.ptr(eigvec), _C.ptr(eigval), _C.ptr(ws),
                                      ws.numel(), _C.stream(m.device))
        _C.check(rc, "ddw_psd_sqrt_fwd")
        ctx.save_for_backward(eigvec, eigval)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        eigvec, eigval = ctx.saved_tensors
        lib = _C.lib()
        B, N, _ = eigvec.shape
        dt = _C.dtype_code(eigvec)
        g = _C.aligned(grad_out)
        d_m = torch.empty_like(eigvec)
        ws = _workspace(lib.ddw_covariance_workspace_bytes(B, 2, N, dt, 1), eigvec.device)
        with _C.device_guard(eigvec):
            rc = lib.ddw_psd_sqrt_bwd(_C.ptr(g), _C.ptr(eigvec), _C.ptr(eigval), B, N, dt, _C.ptr(d_m), _C.ptr(ws),
                                      ws.numel(), _C.stream(eigvec.device))
        _C.check(rc, "ddw_psd_sqrt_bwd")
        return d_m


class CovarianceMatrix(nn.Module):
    """Covariance matrix or its square root.

    Drop-in for ``deepdow.layers.CovarianceMatrix`` (misc.py:33-201).

    Parameters
    ----------
    sqrt : bool
        If True, then returning the square root.
    shrinkage_strategy : None or {'diagonal', 'identity', 'scaled_identity'}
        Strategy of combining the sample covariance matrix with some more stable matrix.
    shrinkage_coef : float or None
        Weight of the sample covariance in the convex combination; if None it has to be given per
        sample in ``forward``.
    """

    def __init__(self, sqrt=True, shrinkage_strategy="diagonal", shrinkage_coef=0.5):
        super().__init__()
        self.sqrt = sqrt
        if shrinkage_strategy is not None:
            if shrinkage_strategy not in {"diagonal", "identity", "scaled_identity"}:
                raise ValueError("Unrecognized shrinkage strategy {}".format(shrinkage_strategy))
        self.shrinkage_strategy = shrinkage_strategy
        self.shrinkage_coef = shrinkage_coef

    def forward(self, x, shrinkage_coef=None):
        """x (n_samples, dim, n_assets); shrinkage_coef None or (n_samples,).  Returns (n_samples, n_assets, n_assets)."""
        n_samples = x.shape[0]
        dtype, device = x.dtype, x.device
        if not ((shrinkage_coef is None) ^ (self.shrinkage_coef is None)):
            raise ValueError("Not clear which shrinkage coefficient to use")
        if shrinkage_coef is not None:
            shrinkage_coef_ = shrinkage_coef
        else:
            shrinkage_coef_ = self.shrinkage_coef * torch.ones(n_samples, dtype=dtype, device=device)
        if x.shape[1] == 1:
            raise ZeroDivisionError("float division by zero")  # misc.py:143: 1.0 / (dim - 1)
        coef = None if self.shrinkage_strategy is None else shrinkage_coef_.reshape(-1)
        return _CovarianceFn.apply(x, coef, _C.SHRINKAGE[self.shrinkage_strategy], bool(self.sqrt))

    @staticmethod
    def compute_covariance(m, shrinkage_strategy=None, shrinkage_coef=0.5):
        """Covariance of a single sample; m has shape (n_assets, n_channels) as in misc.py:121-166."""
        coef = torch.as_tensor(shrinkage_coef, dtype=m.dtype, device=m.device).reshape(1)
        x = m.t().unsqueeze(0)
        if x.shape[1] == 1:
            raise ZeroDivisionError("float division by zero")
        coef = None if shrinkage_strategy is None else coef
        return _CovarianceFn.apply(x, coef, _C.SHRINKAGE[shrinkage_strategy], False)[0]

    @staticmethod
    def compute_sqrt(m):
        """Square root of a single positive semi-definite matrix (n_assets, n_assets), misc.py:168-201."""
        return _PsdSqrtFn.apply(m.unsqueeze(0))[0]


# ------------------------------------------------------------------------------------------------
# Fused CovarianceMatrix(sqrt=False) -> NumericalMarkowitz (BachelierNet's chain)
# ------------------------------------------------------------------------------------------------
class _CovMarkowitzFn(torch.autograd.Function):
    """x (B, L, N) -> covariance in shared memory -> solver: the (B, N, N) matrix never exists in HBM in the forward.  The
    backward recomputes it (one covariance launch) and chains the two ordinary backward kernels."""

    @staticmethod
    def forward(ctx, x, coef, rets, gamma_sqrt, alpha, strategy, max_weight, check_status, watch):
        _C.require_cuda(x, coef, rets, gamma_sqrt, alpha)
        lib = _C.lib()
        B, L, N = x.shape
        dt = _C.dtype_code(x)
        x_c, rets_c = _C.aligned(x), _C.aligned(rets)
        coef_c = None if coef is None else _C.aligned(coef.to(x.dtype))
        gam_c, alp_c = _C.aligned(gamma_sqrt), _C.aligned(alpha)
        w = torch.empty_like(rets_c)
        state = torch.empty((B, N), dtype=torch.int8, device=x.device)
        status = torch.empty((B,), dtype=torch.int32, device=x.device)
        ws = _workspace(_C.markowitz_workspace_bytes(B, N, dt), x.device)
        saved = _workspace(lib.ddw_markowitz_saved_bytes(B, N, dt), x.device) if any(ctx.needs_input_grad[:5]) else None
        # address space for the rare portfolio that needs the interior-point fallback (which reads covmat_sqrt from HBM):
        # never initialised, written only by such a portfolio
        spill = torch.empty((B, N, N), dtype=x.dtype, device=x.device)
        with _C.device_guard(x):
            rc = lib.ddw_covariance_markowitz_fwd(_C.ptr(x_c), _C.ptr(coef_c), strategy, L, _C.ptr(rets_c), _C.ptr(gam_c), _C.ptr(alp_c),
                                                  float(max_weight), B, N, dt, _C.ptr(w), _C.ptr(state), _C.ptr(status), _C.ptr(ws),
                                                  ws.numel(), _C.ptr(saved), 0 if saved is None else saved.numel(), _C.ptr(spill),
                                                  _C.stream(x.device))
        if rc == -2:
            raise _FusedUnsupported()
        _C.check(rc, "ddw_covariance_markowitz_fwd")
        what = f"batch of {B}, n_assets={N}, max_weight={max_weight} (fused covariance)"
        counter = ws[8:12].view(torch.int32)
        if check_status is True:
            bad = int(counter)
            if bad:
                raise SolverError(f"{bad} portfolio problems could not be solved ({what})")
        elif check_status == "lazy":
            with _C.device_guard(x):
                watch.push(counter, what)
        ctx.save_for_backward(x_c, coef_c, rets_c, gam_c, alp_c, w, state, saved)
        ctx.args = (strategy, float(max_weight))
        ctx.mark_non_differentiable(status)
        return w, status

    @staticmethod
    def backward(ctx, grad_w, _grad_status):
        x_c, coef_c, rets_c, gam_c, alp_c, w, state, saved = ctx.saved_tensors
        strategy, max_weight = ctx.args
        lib = _C.lib()
        B, L, N = x_c.shape
        dt = _C.dtype_code(x_c)
        dev = x_c.device
        cov = torch.empty((B, N, N), dtype=x_c.dtype, device=dev)
        cws = _workspace(lib.ddw_covariance_workspace_bytes(B, L, N, dt, 0), dev)
        ws = _workspace(_C.markowitz_workspace_bytes(B, N, dt), dev)
        g = _C.aligned(grad_w)
        d_rets, d_cov, d_alpha, d_gamma = torch.empty_like(w), torch.empty_like(cov), torch.empty_like(alp_c), torch.empty_like(gam_c)
        d_x = torch.empty_like(x_c)
        d_coef = torch.empty_like(coef_c) if (coef_c is not None and ctx.needs_input_grad[1]) else None
        with _C.device_guard(x_c):
            st = _C.stream(dev)
            rc = lib.ddw_covariance_fwd(_C.ptr(x_c), _C.ptr(coef_c), strategy, 0, B, L, N, dt, _C.ptr(cov), None, None, None,
                                        _C.ptr(cws), cws.numel(), st)
            _C.check(rc, "ddw_covariance_fwd")
            if saved is None:
                rc = lib.ddw_markowitz_bwd(_C.ptr(g), _C.ptr(w), _C.ptr(state), _C.ptr(cov), _C.ptr(gam_c), _C.ptr(alp_c), max_weight, B, N,
                                           dt, _C.ptr(d_rets), _C.ptr(d_cov), _C.ptr(d_gamma), _C.ptr(d_alpha), _C.ptr(ws), ws.numel(), st)
            else:
                rc = lib.ddw_markowitz_bwd_saved(_C.ptr(g), _C.ptr(w), _C.ptr(state), _C.ptr(cov), _C.ptr(gam_c), _C.ptr(alp_c), max_weight,
                                                 B, N, dt, _C.ptr(d_rets), _C.ptr(d_cov), _C.ptr(d_gamma), _C.ptr(d_alpha), _C.ptr(ws),
                                                 ws.numel(), _C.ptr(saved), saved.numel(), st)
            _C.check(rc, "ddw_markowitz_bwd")
            rc = lib.ddw_covariance_bwd(_C.ptr(d_cov), _C.ptr(x_c), _C.ptr(coef_c), strategy, 0, None, None, B, L, N, dt, _C.ptr(d_x),
                                        _C.ptr(d_coef), _C.ptr(cws), cws.numel(), st)
            _C.check(rc, "ddw_covariance_bwd")
        return d_x, d_coef, d_rets, d_gamma, d_alpha, None, None, None, None


class _FusedUnsupported(Exception):
    """The fused kernel cannot take this shape: compose the two layers."""


def covariance_markowitz(x, rets, gamma_sqrt, alpha, covariance_layer, allocator, shrinkage_coef=None):
    """``allocator(rets, covariance_layer(x), gamma_sqrt, alpha)`` for a ``CovarianceMatrix(sqrt=False)`` and a
    ``NumericalMarkowitz`` -- BachelierNet's chain (deepdow/nn.py:170-171,189-191) -- through the fused kernel
    ``ddw_covariance_markowitz_fwd``: the (n_samples, n_assets, n_assets) matrix is formed in shared memory and never written
    to HBM in the forward.  Same results as the composition (the covariance tile is the same code); shapes the fused kernel
    cannot take (n_assets > 128, long lookbacks) run the composition."""
    if covariance_layer.sqrt or x.shape[1] == 1 or not x.is_cuda:
        return allocator(rets, covariance_layer(x, shrinkage_coef), gamma_sqrt, alpha)
    n_samples = x.shape[0]
    if not ((shrinkage_coef is None) ^ (covariance_layer.shrinkage_coef is None)):
        raise ValueError("Not clear which shrinkage coefficient to use")
    coef = None
    if covariance_layer.shrinkage_strategy is not None:
        coef = (shrinkage_coef if shrinkage_coef is not None
                else covariance_layer.shrinkage_coef * torch.ones(n_samples, dtype=x.dtype, device=x.device)).reshape(-1)
    dtype = rets.dtype
    if allocator.check_status:
        allocator._watch.poll()
    try:
        w, status = _CovMarkowitzFn.apply(x.to(dtype), coef, rets, gamma_sqrt.reshape(-1).to(dtype), alpha.reshape(-1).to(dtype),
                                          _C.SHRINKAGE[covariance_layer.shrinkage_strategy], allocator.max_weight,
                                          allocator.check_status, allocator._watch)
    except _FusedUnsupported:
        return allocator(rets, covariance_layer(x, shrinkage_coef), gamma_sqrt, alpha)
    allocator.last_status = status
    return w


# ------------------------------------------------------------------------------------------------
# Resample: the bootstrap meta-allocator on top of the two layers above
# ------------------------------------------------------------------------------------------------
class Resample(nn.Module):
    """Meta allocator that bootstraps the input expected returns and covariance matrix.

    The ``NumericalMarkowitz`` arm of ``deepdow.layers.Resample`` (allocate.py:276-411): view ``(rets, matrix)`` as the
    parameters of a multivariate normal, then ``n_portfolios`` times draw ``n_draws`` samples (reparameterised, so gradients
    reach ``rets`` and ``matrix``), re-estimate the expected returns (mean of the draws) and the square root of the covariance
    (``CovarianceMatrix(sqrt=True)`` of the draws) and allocate; the portfolios are averaged.  Same constructor, same
    keyword-only ``forward(matrix, rets, gamma=..., alpha=...)``, same seeding (``torch.manual_seed(random_state)``).
    The reference's other arms (``AnalyticalMarkowitz``, ``NCO``) are not part of this package: they raise ``TypeError``
    here exactly like any other unsupported allocator does there (allocate.py:323-328).
    """

    def __init__(self, allocator, n_draws=None, n_portfolios=5, sqrt=False, random_state=None):
        super().__init__()
        if not isinstance(allocator, NumericalMarkowitz):
            raise TypeError("Unsupported type of allocator: {}".format(type(allocator)))
        self.allocator = allocator
        self.sqrt = sqrt
        self.n_draws = n_draws
        self.n_portfolios = n_portfolios
        self.random_state = random_state
        self.uses_sqrt = True          # NumericalMarkowitz consumes the square root (allocate.py:336-342)
        self.covariance_layer = CovarianceMatrix(sqrt=self.uses_sqrt)

    def forward(self, matrix, rets=None, **kwargs):
        """matrix (n_samples, n_assets, n_assets): the covariance, or its square root when ``sqrt=True``; rets
        (n_samples, n_assets); ``gamma`` and ``alpha`` (n_samples,) as keyword arguments.  Returns (n_samples, n_assets)."""
        if rets is None:
            raise ValueError("NumericalMarkowitz needs expected returns: pass rets")   # the reference fails inside cvxpylayers
        if self.random_state is not None:
            torch.manual_seed(self.random_state)
        n_assets = matrix.shape[1]
        n_draws = self.n_draws or n_assets
        covmat = matrix @ matrix if self.sqrt else matrix
        dist = torch.distributions.MultivariateNormal(loc=rets, covariance_matrix=covmat)
        total = None
        for _ in range(self.n_portfolios):
            draws = dist.rsample((n_draws,))                                   # (n_draws, n_samples, n_assets)
            covmat_ = self.covariance_layer(draws.permute(1, 0, 2))            # the layer takes (n_samples, n_draws, n_assets)
            portfolio = self.allocator(draws.mean(dim=0), covmat_, kwargs["gamma"], kwargs["alpha"])
            total = portfolio if total is None else total + portfolio
        return total / self.n_portfolios

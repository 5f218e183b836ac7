"""ctypes binding of libddw_b200.so (C ABI declared in include/ddw.h).

The library is loaded lazily so that modules holding these layers stay picklable
(``torch.save(network)`` in deepdow/callbacks.py:486) and importable on a CPU-only box.
There is NO CPU fallback: every call needs CUDA tensors and the built library, and fails loudly
otherwise.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DDW_LIB_PATH") or os.path.join(_HERE, "libddw_b200.so")   # override: A/B builds (tools/ab_variants.sh)

DDW_F32, DDW_F64 = 0, 1
SHRINKAGE = {None: 0, "diagonal": 1, "identity": 2, "scaled_identity": 3}
STATUS_FALLBACK, STATUS_REGULARIZED, STATUS_INEXACT, STATUS_INFEASIBLE = 1, 2, 4, 8

_lib = None

_vp, _i64, _i32, _dbl, _sz = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_double, ctypes.c_size_t

# name -> (restype, argtypes); must mirror include/ddw.h exactly (tests/test_abi.py checks the names)
SIGNATURES = {
    "ddw_version": (_i32, []),
    "ddw_last_error": (ctypes.c_char_p, []),
    "ddw_markowitz_workspace_bytes": (_sz, [_i64, _i64, _i32]),
    "ddw_markowitz_fwd": (_i32, [_vp, _vp, _vp, _vp, _dbl, _i64, _i64, _i32, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ddw_markowitz_bwd": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _dbl, _i64, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ddw_markowitz_saved_bytes": (_sz, [_i64, _i64, _i32]),
    "ddw_markowitz_fwd_saved": (_i32, [_vp, _vp, _vp, _vp, _dbl, _i64, _i64, _i32, _vp, _vp, _vp, _vp, _sz, _vp, _sz, _vp]),
    "ddw_markowitz_bwd_saved": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _dbl, _i64, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _sz, _vp, _sz,
                                       _vp]),
    "ddw_covariance_markowitz_fwd": (_i32, [_vp, _vp, _i32, _i64, _vp, _vp, _vp, _dbl, _i64, _i64, _i32, _vp, _vp, _vp, _vp, _sz, _vp, _sz,
                                            _vp, _vp]),
    "ddw_markowitz_fwd_chunk": (_i32, [_vp, _vp, _vp, _vp, _dbl, _i64, _i64, _i32, _i64, _i64, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ddw_markowitz_bwd_chunk": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _dbl, _i64, _i64, _i32, _i64, _i64, _vp, _vp, _vp, _i32,
                                       _vp, _vp, _sz, _vp]),
    "ddw_host_pipeline_create": (_i32, [ctypes.POINTER(ctypes.c_void_p)]),
    "ddw_host_pipeline_destroy": (None, [_vp]),
    "ddw_host_pipeline_graph_launches": (_i32, [_vp]),
    "ddw_markowitz_host_scratch_bytes": (_sz, [_i64, _i64, _i32, _i64]),
    "ddw_markowitz_host_step": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _dbl, _i64, _i64, _i32, _i64, _vp, _vp, _vp, _vp, _vp, _vp,
                                       _vp, _sz, _vp]),
    "ddw_markowitz_host_step_factored": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _dbl, _i64, _i64, _i32, _i64, _vp, _vp, _vp, _vp, _vp, _vp,
                                                _vp, _sz, _vp]),
    "ddw_risk_budgeting_workspace_bytes": (_sz, [_i64, _i64, _i32]),
    "ddw_risk_budgeting_fwd": (_i32, [_vp, _vp, _dbl, _i64, _i64, _i32, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ddw_risk_budgeting_bwd": (_i32, [_vp, _vp, _vp, _vp, _vp, _dbl, _i64, _i64, _i32, _vp, _vp, _vp, _sz, _vp]),
    "ddw_capped_simplex_fwd": (_i32, [_vp, _vp, _dbl, _dbl, _i64, _i64, _i32, _vp, _vp]),
    "ddw_capped_simplex_bwd": (_i32, [_vp, _vp, _vp, _vp, _dbl, _dbl, _i64, _i64, _i32, _vp, _vp, _vp]),
    "ddw_capped_argmax_fwd": (_i32, [_vp, _dbl, _i64, _i64, _i32, _vp, _vp]),
    "ddw_capped_softmax_fwd": (_i32, [_vp, _vp, _dbl, _dbl, _i64, _i64, _i32, _vp, _vp]),
    "ddw_capped_softmax_bwd": (_i32, [_vp, _vp, _vp, _vp, _dbl, _dbl, _i64, _i64, _i32, _vp, _vp, _vp]),
    "ddw_covariance_workspace_bytes": (_sz, [_i64, _i64, _i64, _i32, _i32]),
    "ddw_covariance_fwd": (_i32, [_vp, _vp, _i32, _i32, _i64, _i64, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ddw_covariance_bwd": (_i32, [_vp, _vp, _vp, _i32, _i32, _vp, _vp, _i64, _i64, _i64, _i32, _vp, _vp, _vp, _sz, _vp]),
    "ddw_psd_sqrt_fwd": (_i32, [_vp, _i64, _i64, _i32, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ddw_psd_sqrt_bwd": (_i32, [_vp, _vp, _vp, _i64, _i64, _i32, _vp, _vp, _sz, _vp]),
}


class ExtensionError(RuntimeError):
    """libddw_b200.so is missing or a call into it failed."""


def lib():
    """Load (once) and return the shared library; raise ExtensionError if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ExtensionError(
                f"{LIB_PATH} not found: build it with `python -m deepdow_b200.build` "
                "(there is no CPU or pure-PyTorch fallback)")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().ddw_last_error().decode()
        raise ExtensionError(f"{what} failed with code {rc}: {msg}")


def dtype_code(t):
    if t.dtype == torch.float32:
        return DDW_F32
    if t.dtype == torch.float64:
        return DDW_F64
    raise TypeError(f"deepdow_b200 kernels take float32 or float64 tensors, got {t.dtype}")


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise ExtensionError(
                "deepdow_b200 layers run on CUDA tensors only (B200 kernels, no CPU fallback); "
                f"got a tensor on {t.device}")


class _NoGuard:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


_NO_GUARD = _NoGuard()


def device_guard(t):
    """Context manager that makes `t`'s GPU the current CUDA device for the launch (kernels go to the current
    device); free when it already is -- the common case pays one integer comparison."""
    if t.device.index is None or t.device.index == torch.cuda.current_device():
        return _NO_GUARD
    return torch.cuda.device(t.device)


def aligned(t):
    """Contiguous and 16-byte aligned (the kernels use 128-bit loads/stores); copies only if needed."""
    if t is None:
        return None
    t = t.contiguous()
    return t.clone() if t.data_ptr() % 16 else t


def ptr(t):
    """Device (or host) address of a tensor for a ``c_void_p`` argument: a plain int converts for free; None is NULL."""
    return None if t is None else t.data_ptr()


_ws_bytes_cache = {}


def markowitz_workspace_bytes(B, N, dt):
    """ddw_markowitz_workspace_bytes, memoised (it is a pure function of the shape; the call costs ~2 us of ctypes)."""
    key = (B, N, dt)
    v = _ws_bytes_cache.get(key)
    if v is None:
        v = _ws_bytes_cache[key] = int(lib().ddw_markowitz_workspace_bytes(B, N, dt))
    return v


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def stream(device=None):
    """cudaStream_t of torch's current stream on `device` (default: the current device) as a c_void_p.

    ``torch.cuda.current_stream()`` builds a Python Stream object (about 20 us per call, measured with
    tools/host_profile.py); the raw accessor is the same lookup without the object."""
    if _raw_stream is not None:
        index = torch.cuda.current_device() if device is None else torch.device(device).index
        return ctypes.c_void_p(_raw_stream(index if index is not None else torch.cuda.current_device()))
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)

"""GPU parity of the chunked C-ABI entry points and of the host-buffer pipeline
(ddw_markowitz_fwd_chunk / ddw_markowitz_bwd_chunk / ddw_markowitz_host_step, include/ddw.h):
a batch solved in chunks, or streamed from host buffers, must give bit-identical weights and the same
gradients as the one-launch device path, including the reference's gamma tiling
(deepdow/layers/allocate.py:268-270), which indexes gamma_sqrt by the flat element index of the WHOLE batch.
"""
import numpy as np
import pytest
import torch

import oracle
from deepdow_b200 import NumericalMarkowitz, SolverError, _C
from deepdow_b200.host import MarkowitzHostPipeline

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def problems(B, N, seed, gamma="rand", dtype=np.float32):
    rng = np.random.default_rng(seed)
    R = rng.standard_normal((B, N, N))
    C = R @ R.transpose(0, 2, 1) / N + 0.1 * np.eye(N)
    rets = rng.standard_normal((B, N)) * 0.5
    g = rng.uniform(0.5, 2.0, B) if gamma == "rand" else np.ones(B)
    a = rng.uniform(0.5, 1.5, B) * rng.choice([-1.0, 1.0], B)
    G = rng.standard_normal((B, N))
    return [x.astype(dtype) for x in (rets, C, g, a, G)]


def device_reference(rets, C, g, a, G, u):
    """One launch of ddw_markowitz_fwd + ddw_markowitz_bwd over the whole batch (the entry points the chunked calls and the
    host pipeline are built from; the nn.Module itself takes the saved-factor pair, whose backward rounds differently)."""
    lib = _C.lib()
    tr, tC, tg, ta, tG = (torch.tensor(x, device=DEV) for x in (rets, C, g, a, G))
    B, N = tr.shape
    dt = _C.dtype_code(tr)
    w = torch.empty_like(tr); state = torch.empty(B, N, dtype=torch.int8, device=DEV); status = torch.empty(B, dtype=torch.int32, device=DEV)
    ws = torch.empty(lib.ddw_markowitz_workspace_bytes(B, N, dt), dtype=torch.uint8, device=DEV)
    d_r, d_C, d_g, d_a = torch.empty_like(tr), torch.empty_like(tC), torch.empty_like(tg), torch.empty_like(ta)
    _C.check(lib.ddw_markowitz_fwd(_C.ptr(tr), _C.ptr(tC), _C.ptr(tg), _C.ptr(ta), u, B, N, dt, _C.ptr(w), _C.ptr(state), _C.ptr(status),
                                   _C.ptr(ws), ws.numel(), _C.stream()), "fwd")
    _C.check(lib.ddw_markowitz_bwd(_C.ptr(tG), _C.ptr(w), _C.ptr(state), _C.ptr(tC), _C.ptr(tg), _C.ptr(ta), u, B, N, dt, _C.ptr(d_r),
                                   _C.ptr(d_C), _C.ptr(d_g), _C.ptr(d_a), _C.ptr(ws), ws.numel(), _C.stream()), "bwd")
    torch.cuda.synchronize()
    return [x.cpu() for x in (w, d_r, d_C, d_g, d_a)]


@pytest.mark.parametrize("B,N,u,chunk", [(96, 20, 0.3, 32), (70, 12, 1.0, 32), (50, 100, 0.2, 16), (33, 7, 0.5, 4)])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_chunked_abi_matches_one_launch(B, N, u, chunk, dtype):
    rets, C, g, a, G = problems(B, N, 100 + B + N, dtype=dtype)
    ref = device_reference(rets, C, g, a, G, u)
    lib = _C.lib()
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    d = lambda x: torch.tensor(x, device=DEV)
    trets, tC, tg, ta, tG = d(rets), d(C), d(g), d(a), d(G)
    dt = _C.dtype_code(trets)
    w = torch.empty(B, N, dtype=tdt, device=DEV)
    state = torch.empty(B, N, dtype=torch.int8, device=DEV)
    status = torch.empty(B, dtype=torch.int32, device=DEV)
    d_rets, d_C, d_g, d_a = torch.empty_like(trets), torch.empty_like(tC), torch.zeros_like(tg), torch.empty_like(ta)
    for b0 in range(0, B, chunk):
        c = min(chunk, B - b0)
        ws = torch.empty(lib.ddw_markowitz_workspace_bytes(c, N, dt), dtype=torch.uint8, device=DEV)
        sl = slice(b0, b0 + c)
        rc = lib.ddw_markowitz_fwd_chunk(_C.ptr(trets[sl]), _C.ptr(tC[sl]), _C.ptr(tg), _C.ptr(ta[sl]), u, c, N, dt, B, b0,
                                         _C.ptr(w[sl]), _C.ptr(state[sl]), _C.ptr(status[sl]), _C.ptr(ws), ws.numel(),
                                         _C.stream())
        _C.check(rc, "fwd_chunk")
        rc = lib.ddw_markowitz_bwd_chunk(_C.ptr(tG[sl]), _C.ptr(w[sl]), _C.ptr(state[sl]), _C.ptr(tC[sl]), _C.ptr(tg),
                                         _C.ptr(ta[sl]), u, c, N, dt, B, b0, _C.ptr(d_rets[sl]), _C.ptr(d_C[sl]),
                                         _C.ptr(d_g), 1, _C.ptr(d_a[sl]), _C.ptr(ws), ws.numel(), _C.stream())
        _C.check(rc, "bwd_chunk")
    torch.cuda.synchronize()
    # the chunk sizes above are multiples of 4 so that every slice passed to the ABI stays 16-byte aligned
    assert torch.equal(w.cpu(), ref[0])
    assert torch.equal(d_rets.cpu(), ref[1])
    assert torch.equal(d_C.cpu(), ref[2])
    assert torch.equal(d_a.cpu(), ref[4])
    tol = 1e-5 if dtype == np.float32 else 1e-12
    assert (d_g.cpu() - ref[3]).abs().max() <= tol * max(1.0, float(ref[3].abs().max()))   # atomics: order differs


@pytest.mark.parametrize("B,N,u,chunk", [(100, 20, 0.3, 32), (64, 100, 0.2, 16), (20, 12, 1.0, 64), (45, 8, 0.5, 7)])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_host_pipeline_matches_device_path_and_oracle(B, N, u, chunk, dtype):
    rets, C, g, a, G = problems(B, N, 7 * B + N, dtype=dtype)
    ref = device_reference(rets, C, g, a, G, u)
    pipe = MarkowitzHostPipeline(N, u, chunk=chunk, device=DEV)
    host = [torch.tensor(x).pin_memory() for x in (rets, C, g, a, G)]
    out = pipe.step(*host)
    assert torch.equal(out.w, ref[0]) and int(out.status.ne(0).sum()) == 0
    out = pipe.step(*host[:4])                      # forward only
    assert torch.equal(out.w, ref[0]) and out.d_rets is None
    out = None
    for _ in range(3):                              # repeated steps recycle slots and output buffers
        out = pipe.step(*host, out=out)
    assert pipe.graph_launches >= 2                 # same buffers again: captured once, then replayed
    assert torch.equal(out.w, ref[0])
    assert torch.equal(out.d_rets, ref[1]) and torch.equal(out.d_covmat_sqrt, ref[2]) and torch.equal(out.d_alpha, ref[4])
    tol = 1e-5 if dtype == np.float32 else 1e-12
    assert (out.d_gamma_sqrt - ref[3]).abs().max() <= tol * max(1.0, float(ref[3].abs().max()))
    # and against the CPU oracle
    w_ref, st_ref, kkt, _ = oracle.markowitz_fwd(rets, C, g, a, u)
    d_ref = oracle.markowitz_bwd(G, w_ref, st_ref, C, g, a, u)
    wtol, gtol = (1e-5, 1e-4) if dtype == np.float32 else (1e-9, 1e-8)
    assert np.abs(out.w.double().numpy() - w_ref).max() < wtol
    for got, want in zip(out[1:5], d_ref):
        assert np.abs(got.double().numpy() - want).max() <= gtol * max(np.abs(want).max(), 1e-30)
    pipe.close()


def test_host_pipeline_errors():
    pipe = MarkowitzHostPipeline(10, 0.05, device=DEV)
    z = torch.zeros
    with pytest.raises(SolverError):
        pipe.step(z(4, 10), z(4, 10, 10), z(4), z(4))
    pipe = MarkowitzHostPipeline(10, 0.5, device=DEV)
    with pytest.raises(ValueError):
        pipe.step(z(4, 9), z(4, 9, 9), z(4), z(4))
    with pytest.raises(ValueError):
        pipe.step(z(4, 10, device=DEV), z(4, 10, 10), z(4), z(4))
    # pageable (non-pinned) host memory is accepted
    rets, C, g, a, G = problems(8, 10, 3)
    out = pipe.step(*[torch.tensor(x) for x in (rets, C, g, a, G)])
    assert abs(float(out.w.sum()) - 8.0) < 1e-4


def test_module_streams_cpu_tensors_when_opted_in():
    """NumericalMarkowitz.host_device: CPU tensors in, CPU tensors out (same dtype/device as the inputs, like the
    reference, tests/test_layers.py:447-448), gradients equal to the device path; default stays "CUDA tensors only"."""
    from deepdow_b200._C import ExtensionError
    B, N, u = 40, 16, 0.4
    rets, C, g, a, G = problems(B, N, 11)
    ref = device_reference(rets, C, g, a, G, u)
    layer = NumericalMarkowitz(N, u)
    cpu = [torch.tensor(x, requires_grad=True) for x in (rets, C, g, a)]
    with pytest.raises(ExtensionError):
        layer(*cpu)
    layer.host_device = DEV
    w = layer(*cpu)
    assert w.device.type == "cpu" and w.dtype == torch.float32 and torch.equal(w.detach(), ref[0])
    (w * torch.tensor(G)).sum().backward()
    assert torch.equal(cpu[0].grad, ref[1]) and torch.equal(cpu[1].grad, ref[2]) and torch.equal(cpu[3].grad, ref[4])
    assert (cpu[2].grad - ref[3]).abs().max() <= 1e-5 * max(1.0, float(ref[3].abs().max()))
    import pickle
    layer2 = pickle.loads(pickle.dumps(layer))
    assert torch.equal(layer2(*[t.detach() for t in cpu]), ref[0])


@pytest.mark.parametrize("B,N,u,chunk,dense", [(96, 20, 0.3, 32, False), (50, 100, 0.2, 16, False), (40, 100, 0.2, 16, True)])
def test_host_pipeline_factored_gradient(B, N, u, chunk, dense):
    """ddw_markowitz_host_step_factored: the rank-2 gradient of covmat_sqrt comes back as its factor vectors (2 N numbers per
    portfolio); expanded on the host (with the reference's gamma tiling) it equals the dense gradient of the regular step, and
    every other output is bit-identical."""
    from deepdow_b200.host import expand_factored
    rets, C, g, a, G = problems(B, N, 300 + B + N)
    if dense:
        rets = rets * 0.04
    host = [torch.tensor(x).pin_memory() for x in (rets, C, g, a, G)]
    pipe = MarkowitzHostPipeline(N, u, chunk=chunk, device=DEV)
    full = pipe.step(*host)
    fac = pipe.step(*host, factored=True)
    fac2 = pipe.step(*host, factored=True, out=fac)           # second identical request: served by the cached CUDA graph
    assert fac.factors.shape == (B, 2, N)
    for name in ("w", "d_rets", "d_alpha", "status"):
        assert torch.equal(getattr(full, name), getattr(fac2, name)), name
    assert (full.d_gamma_sqrt - fac2.d_gamma_sqrt).abs().max() <= 1e-5 * max(1.0, float(full.d_gamma_sqrt.abs().max()))   # atomics
    d_cov = expand_factored(fac2, host[2])
    assert (d_cov - full.d_covmat_sqrt).abs().max() <= 1e-6 * max(1.0, float(full.d_covmat_sqrt.abs().max()))
    pipe.close()

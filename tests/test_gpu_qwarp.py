"""GPU parity of the tensor-core Gram + warp-per-portfolio NumericalMarkowitz forward (csrc/markowitz_qwarp.cu), the float32 path
for n_assets <= 128: its own corner cases on top of the shapes of test_gpu_parity.py (which all run through it by default)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import oracle
from deepdow_b200 import NumericalMarkowitz

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(autouse=True)
def _tensor_core_path_for_small_batches(monkeypatch):
    """By default batches below 1024 portfolios take the one-CTA-per-portfolio kernel (faster there); these tests are about the
    tensor-core path, at sizes the oracle finishes in seconds."""
    monkeypatch.setenv("DDW_QWARP_MIN_BATCH", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _problem(B, N, seed, rscale=0.5, offdiag=1.0):
    rng = np.random.default_rng(seed)
    R = rng.standard_normal((B, N, N)).astype(np.float32)
    C = (offdiag * R @ R.transpose(0, 2, 1) / N + 0.1 * np.eye(N, dtype=np.float32)).astype(np.float32)
    rets = (rng.standard_normal((B, N)) * rscale).astype(np.float32)
    return rng, rets, C


def _solve(rets, C, g, a, u):
    layer = NumericalMarkowitz(rets.shape[1], u)
    layer.check_status = True
    with torch.no_grad():
        w = layer(*[torch.tensor(x, device=DEV) for x in (rets, C, g, a)])
    return w.double().cpu().numpy(), layer


def _oracle(rets, C, g, a, u):
    return oracle.markowitz_fwd(rets.astype(np.float64), C.astype(np.float64), g.astype(np.float64), a.astype(np.float64), u)[:2]


@pytest.mark.parametrize("B,N,u,rscale", [
    (64, 8, 1.0, 0.5),        # smallest universe of the path, one 16-column MMA tile
    (96, 36, 0.3, 0.5),       # n_assets % 4 == 0 but % 16 != 0: K tail inside a chunk, padded MMA columns
    (64, 37, 0.3, 0.5),       # odd n_assets: 4 N^2 bytes not a multiple of 16 -> no bulk TMA, scalar staging path
    (48, 64, 0.2, 0.5),       # exactly two 32-k chunks
    (40, 128, 0.1, 0.5),      # largest universe: 128 x 128 tile, one raw buffer
    (128, 50, 0.5, 0.15),     # free sets around the 32-row frame: trimmed ones and hand-overs to the dense kernel
    (64, 100, 0.025, 0.5),    # many assets at the cap: free + capped > 32 -> dense kernel
    (64, 100, 0.04, 3.0),     # a few capped assets inside the frame (rows of capped assets are gathered too)
])
def test_weights_and_states_match_the_oracle(B, N, u, rscale):
    rng, rets, C = _problem(B, N, 100 * N + B, rscale)
    g = (0.5 + rng.random(B)).astype(np.float32)
    a = (rng.random(B) * 2 - 1).astype(np.float32)
    w, layer = _solve(rets, C, g, a, u)
    w_ref, st_ref = _oracle(rets, C, g, a, u)
    assert np.abs(w - w_ref).max() < 1e-5                       # north star: weights within 1e-5 abs
    assert int(layer.last_status.ne(0).sum()) == 0
    # supports: an asset the oracle holds strictly inside (0, u) by more than the tolerance is inside here too, and vice versa
    inside_ref = (w_ref > 1e-5) & (w_ref < u - 1e-5)
    inside = (w > 0) & (w < u)
    assert not (inside_ref & ~inside).any()
    assert np.allclose(w.sum(1), 1.0, atol=1e-5)


@pytest.mark.parametrize("Bmod_like", ["uniform", "per_sample", "odd_batch"])
def test_gamma_tiling_variants(Bmod_like):
    """gamma_sqrt the same for the whole batch (scaled in the epilogue), different per sample with B % 4 == 0 (16-byte gamma loads)
    and with an odd batch (scalar gamma loads): the reference's repeat/view tiling (allocate.py:268-270) in all three."""
    B = 33 if Bmod_like == "odd_batch" else 32
    N, u = 40, 0.25
    rng, rets, C = _problem(B, N, 7)
    g = np.full(B, 1.3, np.float32) if Bmod_like == "uniform" else (0.5 + rng.random(B)).astype(np.float32)
    a = np.full(B, 0.4, np.float32)
    w, _ = _solve(rets, C, g, a, u)
    w_ref, _ = _oracle(rets, C, g, a, u)
    assert np.abs(w - w_ref).max() < 1e-5


def test_tensor_core_path_equals_the_cta_path():
    """A/B inside one test: the same batch through the tensor-core forward (default) and through the one-CTA-per-portfolio
    float32 kernel (DDW_NO_QWARP=1, read once per process -> a child process): weights equal to rounding, states equal."""
    code = r'''
import sys, numpy as np, torch
sys.path.insert(0, %r)
from deepdow_b200 import NumericalMarkowitz
rng = np.random.default_rng(5)
B, N, u = 512, 100, 0.2
R = rng.standard_normal((B, N, N)).astype(np.float32)
C = (R @ R.transpose(0, 2, 1) / N + 0.1 * np.eye(N, dtype=np.float32)).astype(np.float32)
rets = (rng.standard_normal((B, N)) * 0.5).astype(np.float32)
layer = NumericalMarkowitz(N, u); layer.check_status = True
with torch.no_grad():
    w = layer(torch.tensor(rets, device="cuda:0"), torch.tensor(C, device="cuda:0"), torch.ones(B, device="cuda:0"), torch.ones(B, device="cuda:0"))
np.save(sys.argv[1], w.cpu().numpy())
''' % ROOT
    import tempfile
    outs = []
    with tempfile.TemporaryDirectory() as tmp:
        for i, env_extra in enumerate(({"DDW_QWARP_MIN_BATCH": "1"}, {"DDW_NO_QWARP": "1"})):
            env = dict(os.environ, **env_extra)
            if i == 0:
                env.pop("DDW_NO_QWARP", None)
            path = os.path.join(tmp, f"w{i}.npy")
            subprocess.run([sys.executable, "-c", code, path], check=True, env=env, timeout=600)
            outs.append(np.load(path))
    assert np.abs(outs[0] - outs[1]).max() < 2e-6
    assert ((outs[0] > 0) == (outs[1] > 0)).all()
    assert ((outs[0] >= 0.2) == (outs[1] >= 0.2)).all()


def test_forward_is_bitwise_repeatable_and_graph_capturable():
    B, N, u = 300, 100, 0.2
    _, rets, C = _problem(B, N, 11)
    args = [torch.tensor(x, device=DEV) for x in (rets, C, np.ones(B, np.float32), np.ones(B, np.float32))]
    layer = NumericalMarkowitz(N, u)
    layer.check_status = False
    with torch.no_grad():
        w0 = layer(*args).clone()
        for _ in range(3):
            assert torch.equal(layer(*args), w0)
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            layer(*args)
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            wg = layer(*args)
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(wg, w0)


def test_non_finite_inputs_raise_and_do_not_touch_their_neighbours():
    """NaN in rets, Inf in covmat_sqrt: the affected portfolios are reported (SolverError, the reference's solver fails loudly
    too), the kernels neither hang nor trap, and every other portfolio of the batch is solved as usual."""
    from deepdow_b200 import SolverError
    B, N, u = 96, 40, 0.3
    rng, rets, C = _problem(B, N, 21)
    g, a = np.ones(B, np.float32), np.ones(B, np.float32)
    w_ok, _ = _solve(rets, C, g, a, u)
    bad_rets = rets.copy(); bad_rets[5, 7] = np.nan
    bad_C = C.copy(); bad_C[11, 3, 3] = np.inf
    for r_, c_, idx in ((bad_rets, C, 5), (rets, bad_C, 11)):
        layer = NumericalMarkowitz(N, u); layer.check_status = True
        with pytest.raises(SolverError):
            layer(*[torch.tensor(x, device=DEV) for x in (r_, c_, g, a)])
        layer2 = NumericalMarkowitz(N, u); layer2.check_status = False
        with torch.no_grad():
            w = layer2(*[torch.tensor(x, device=DEV) for x in (r_, c_, g, a)]).double().cpu().numpy()
        keep = np.arange(B) != idx
        assert np.abs(w[keep] - w_ok[keep]).max() < 1e-6
        assert int(layer2.last_status[idx]) != 0


def test_chunked_entry_point_keeps_the_gamma_tiling():
    """ddw_markowitz_host_step streams the batch in chunks: every chunk runs the tensor-core forward with the batch offset and the
    whole batch's gamma modulus (allocate.py:268-270 tiles gamma over the flattened batch).  Same weights as the resident call."""
    from deepdow_b200.host import MarkowitzHostPipeline
    B, N, u = 48, 36, 0.3
    rng, rets, C = _problem(B, N, 31)
    g = (0.5 + rng.random(B)).astype(np.float32)
    a = (rng.random(B) * 2 - 1).astype(np.float32)
    G = rng.standard_normal((B, N)).astype(np.float32)
    w_res, _ = _solve(rets, C, g, a, u)
    pipe = MarkowitzHostPipeline(N, u, chunk=16, device=DEV)
    out = pipe.step(*[torch.tensor(x).pin_memory() for x in (rets, C, g, a, G)])
    pipe.close()
    assert np.abs(out.w.double().numpy() - w_res).max() < 1e-6
    w_ref, _ = _oracle(rets, C, g, a, u)
    assert np.abs(out.w.double().numpy() - w_ref).max() < 1e-5


def test_even_universe_not_a_multiple_of_four():
    """n_assets = 38: 4 N^2 bytes is a multiple of 16 (bulk TMA) but rows are not groups of four (scalar staging path)."""
    B, N, u = 64, 38, 0.3
    rng, rets, C = _problem(B, N, 41)
    g = (0.5 + rng.random(B)).astype(np.float32)
    a = np.full(B, 0.2, np.float32)
    w, _ = _solve(rets, C, g, a, u)
    w_ref, _ = _oracle(rets, C, g, a, u)
    assert np.abs(w - w_ref).max() < 1e-5


@pytest.mark.parametrize("dense_share", [1.0, 0.5, 0.1])
def test_probe_routes_diversified_batches_to_the_dense_kernel(dense_share, monkeypatch):
    """A batch of 1024 (the default threshold of the tensor-core path) whose portfolios are diversified (small expected returns:
    nearly every asset free), concentrated, or a mix: the probe kernel decides per call whether the Gram runs at all; the weights
    match the oracle in every case and the hand-over counter says which way the batch went."""
    monkeypatch.delenv("DDW_QWARP_MIN_BATCH", raising=False)
    B, N, u = 1024, 48, 0.25
    rng, rets, C = _problem(B, N, 51)
    ndense = int(B * dense_share)
    order = rng.permutation(B)
    rets[order[:ndense]] *= 0.02                      # diversified optimum: ~all 48 assets free, beyond the 32-row frame
    g = np.ones(B, np.float32)
    a = np.full(B, 0.5, np.float32)
    w, layer = _solve(rets, C, g, a, u)
    w_ref, st_ref = _oracle(rets, C, g, a, u)
    assert np.abs(w - w_ref).max() < 1e-5
    handed_over = layer.solver_counters()[1]
    free = (np.asarray(st_ref) == 0).sum(1)
    if dense_share == 1.0:
        assert handed_over == B                       # the probe sent the whole batch to the dense kernel
    else:
        assert handed_over <= int((free > 32).sum())  # only portfolios that cannot fit the frame leave the warp solver

"""CovarianceMatrix at large n_assets (tensor-core Gram + blocked Jacobi, eig_block.cu): accuracy against torch float64 eigh
and GPU time.  Usage: python tools/time_cov_large.py [B] [L] [N]"""
import sys
import torch
sys.path.insert(0, ".")
from deepdow_b200 import CovarianceMatrix, _C

B, L, N = (int(a) for a in (sys.argv[1:4] + ["64", "250", "500"][len(sys.argv) - 1:]))
dev = "cuda:0"
gen = torch.Generator(device=dev).manual_seed(5)
x = torch.randn(B, L, N, generator=gen, device=dev) * 0.01


def timeit(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for strategy in ("diagonal", "scaled_identity", "identity", None):
    for sqrt in (False, True):
        layer = CovarianceMatrix(sqrt=sqrt, shrinkage_strategy=strategy)
        y = layer(x)
        nb = min(B, 8)
        xd = x[:nb].double()
        xc = xd - xd.mean(1, keepdim=True)
        S = xc.transpose(1, 2) @ xc / (L - 1)
        eye = torch.eye(N, device=dev, dtype=torch.float64)
        c = 0.5
        if strategy == "diagonal":
            ref = c * S + (1 - c) * torch.diag_embed(torch.diagonal(S, dim1=1, dim2=2))
        elif strategy == "identity":
            ref = c * S + (1 - c) * eye
        elif strategy == "scaled_identity":
            ref = c * S + (1 - c) * torch.diagonal(S, dim1=1, dim2=2).mean(1)[:, None, None] * eye
        else:
            ref = S
        if sqrt:
            lam, V = torch.linalg.eigh(ref)
            thr = lam.abs().max(1, keepdim=True).values * N * torch.finfo(torch.float32).eps
            r = torch.where(lam.abs() > thr, lam.abs().sqrt(), torch.zeros_like(lam))
            ref = (V * r[:, None, :]) @ V.transpose(1, 2)
        err = ((y[:nb].double() - ref).abs().amax((1, 2)) / ref.abs().amax((1, 2))).max().item()
        ms = timeit(lambda: layer(x))
        print(f"B={B} L={L} N={N} strategy={strategy} sqrt={sqrt}: rel err {err:.2e}  {ms:.2f} ms  ({B / ms * 1e3:.0f} mats/s)", flush=True)

# sweeps used by the blocked Jacobi (through the C ABI directly)
lib = _C.lib()
out = torch.empty(B, N, N, device=dev); sweeps = torch.zeros(B, dtype=torch.int32, device=dev)
coef = torch.full((B,), 0.5, device=dev)
ws = torch.empty(lib.ddw_covariance_workspace_bytes(B, L, N, 0, 1), dtype=torch.uint8, device=dev)
rc = lib.ddw_covariance_fwd(_C.ptr(x), _C.ptr(coef), 1, 1, B, L, N, 0, _C.ptr(out), None, None, _C.ptr(sweeps), _C.ptr(ws), ws.numel(), _C.stream(x.device))
torch.cuda.synchronize()
print("rc", rc, "sweeps min/max", sweeps.min().item(), sweeps.max().item())

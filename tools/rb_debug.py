"""Debug: NumericalRiskBudgeting kernel status / iterations / error vs the oracle over sizes (run on the GPU box)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import oracle
from deepdow_b200 import _C
lib = _C.lib()
for dtype in (torch.float32, torch.float64):
    for (B, N, u) in [(4, 100, 0.02), (4, 128, 0.02), (4, 129, 0.02), (4, 150, 0.02), (4, 150, 1.0), (2, 200, 0.02), (2, 300, 1.0)]:
        rng = np.random.default_rng(B * 13 + N)
        R = rng.random((B, N, N)); A = R @ R / N + np.eye(N)
        b = rng.random((B, N)); b /= b.sum(1, keepdims=True)
        tA = torch.tensor(A, dtype=dtype, device="cuda"); tb = torch.tensor(b, dtype=dtype, device="cuda")
        w = torch.empty_like(tb); st = torch.empty(B, N, dtype=torch.int8, device="cuda"); status = torch.empty(B, dtype=torch.int32, device="cuda")
        dt = _C.dtype_code(tb)
        ws = torch.zeros(lib.ddw_risk_budgeting_workspace_bytes(B, N, dt), dtype=torch.uint8, device="cuda")
        rc = lib.ddw_risk_budgeting_fwd(_C.ptr(tA), _C.ptr(tb), u, B, N, dt, _C.ptr(w), _C.ptr(st), _C.ptr(status), _C.ptr(ws), ws.numel(), _C.stream())
        torch.cuda.synchronize()
        w_ref, st_ref, kkt, _ = oracle.risk_budgeting_fwd(tA.double().cpu().numpy(), tb.double().cpu().numpy(), u)
        s = status.cpu().numpy()
        print(dtype, B, N, u, "rc", rc, "flags", (s & 0xFFFF).tolist(), "iters", (s >> 16).tolist(), "err", np.abs(w.double().cpu().numpy() - w_ref).max(),
              "capped", st.sum(1).tolist(), "ref capped", st_ref.sum(1).tolist(), flush=True)

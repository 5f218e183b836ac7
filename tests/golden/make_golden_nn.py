"""Golden fixtures for the batched callers in deepdow_b200/nn.py.  Run in the DEV container only:

    python tests/golden/make_golden_nn.py

The UNMODIFIED reference modules ``deepdow/layers/transform.py`` (RNN), ``deepdow/layers/collapse.py`` (AttentionCollapse)
and ``deepdow/losses.py`` (SharpeRatio) are loaded by file path from /root/reference (they import only torch), run on seeded
inputs, and their parameters, inputs and outputs written to nn_reference.npz.  /root/reference does not exist on the GPU box.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/deepdow"


def load(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    torch.manual_seed(11)
    out = {}
    transform = load("ref_transform", "layers/transform.py")
    collapse = load("ref_collapse", "layers/collapse.py")
    for tag, kw in [("lstm", dict(cell_type="LSTM", bidirectional=True, n_layers=1)),
                    ("rnn", dict(cell_type="RNN", bidirectional=False, n_layers=2))]:
        ref = transform.RNN(2, hidden_size=8, **kw).double()
        x = torch.randn(3, 2, 5, 4, dtype=torch.float64)
        out[f"{tag}_x"] = x.numpy()
        out[f"{tag}_y"] = ref(x).detach().numpy()
        for k, v in ref.state_dict().items():
            out[f"{tag}_p_{k}"] = v.numpy()
    ref = collapse.AttentionCollapse(6).double()
    x = torch.randn(3, 6, 5, 4, dtype=torch.float64)
    out["att_x"] = x.numpy()
    out["att_y"] = ref(x).detach().numpy()
    for k, v in ref.state_dict().items():
        out[f"att_p_{k}"] = v.numpy()
    # losses.py starts with `from .layers import CovarianceMatrix`, and deepdow/layers/__init__.py imports cvxpy (absent
    # here).  Give it a `deepdow.layers` that holds the reference's own CovarianceMatrix (misc.py, loaded by path); the
    # SharpeRatio path does not touch it.
    pkg = types.ModuleType("deepdow")
    pkg.__path__ = [REF]
    sys.modules["deepdow"] = pkg
    lay = types.ModuleType("deepdow.layers")
    lay.CovarianceMatrix = load("ref_misc", "layers/misc.py").CovarianceMatrix
    sys.modules["deepdow.layers"] = lay
    losses = load("deepdow.losses", "losses.py")
    w = torch.softmax(torch.randn(4, 5, dtype=torch.float64), dim=1)
    y = torch.randn(4, 2, 6, 5, dtype=torch.float64) * 0.02
    out["sharpe_w"] = w.numpy()
    out["sharpe_y"] = y.numpy()
    out["sharpe_loss"] = losses.SharpeRatio()(w, y).numpy()
    np.savez_compressed(os.path.join(HERE, "nn_reference.npz"), **out)
    print("wrote nn_reference.npz with", len(out), "arrays")


if __name__ == "__main__":
    main()

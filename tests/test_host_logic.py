"""CPU-only tests of the host-side mirror of the reference layers: constructor / forward error
behaviour (reference tests/test_layers.py:203-209, 677-690, 780-783), no CPU fallback, picklability,
and the analytical softmax branch (pure torch, as in the reference)."""
import io
import pickle

import pytest
import torch

from deepdow_b200 import (CovarianceMatrix, NumericalMarkowitz, SoftmaxAllocator, SolverError,
                          SparsemaxAllocator)
from deepdow_b200._C import ExtensionError


def test_covariance_construction_errors():
    with pytest.raises(ValueError):
        CovarianceMatrix(shrinkage_strategy="fake")
    with pytest.raises(ValueError):
        CovarianceMatrix(shrinkage_coef=None)(torch.ones(1, 5, 2))
    with pytest.raises(ValueError):
        CovarianceMatrix(shrinkage_coef=0.5)(torch.ones(1, 5, 2), torch.ones(1))
    with pytest.raises(ZeroDivisionError):     # misc.py:143, tests/test_layers.py:221-225
        CovarianceMatrix()(torch.ones(2, 1, 3))
    assert sum(p.numel() for p in CovarianceMatrix().parameters()) == 0


def test_softmax_errors():
    with pytest.raises(ValueError):
        SoftmaxAllocator(formulation="wrong")
    with pytest.raises(ValueError):
        SoftmaxAllocator(formulation="variational", n_assets=None)
    with pytest.raises(ValueError):
        SoftmaxAllocator(formulation="analytical", max_weight=0.3)
    with pytest.raises(ValueError):
        SoftmaxAllocator(formulation="variational", max_weight=0.09, n_assets=10)
    for formulation in ("analytical", "variational"):
        with pytest.raises(ValueError):
            SoftmaxAllocator(temperature=None, formulation=formulation, n_assets=4)(torch.ones(2, 4), temperature=None)
        with pytest.raises(ValueError):
            SoftmaxAllocator(temperature=1, formulation=formulation, n_assets=4)(torch.ones(2, 4), torch.ones(2))


def test_sparsemax_errors():
    with pytest.raises(ValueError):
        SparsemaxAllocator(n_assets=2, max_weight=0.3)
    with pytest.raises(ValueError):
        SparsemaxAllocator(4, temperature=None)(torch.ones(2, 4), temperature=None)


def test_analytical_softmax_is_torch_softmax():
    x = torch.randn(3, 6)
    w = SoftmaxAllocator(temperature=2.0)(x)
    assert torch.allclose(w, torch.softmax(x / 2.0, dim=1))
    w2 = SoftmaxAllocator(temperature=None)(x, 2.0 * torch.ones(3))
    assert torch.allclose(w, w2)


def test_no_cpu_fallback():
    """The product path must fail loudly on CPU tensors instead of silently computing elsewhere."""
    with pytest.raises(ExtensionError):
        NumericalMarkowitz(5)(torch.zeros(2, 5), torch.zeros(2, 5, 5), torch.ones(2), torch.ones(2))
    with pytest.raises(ExtensionError):
        SparsemaxAllocator(5)(torch.zeros(2, 5))
    with pytest.raises(ExtensionError):
        SoftmaxAllocator(formulation="variational", n_assets=5)(torch.zeros(2, 5))
    with pytest.raises(ExtensionError):
        CovarianceMatrix()(torch.zeros(2, 4, 5))


def test_markowitz_shape_checks():
    layer = NumericalMarkowitz(5)
    with pytest.raises(ValueError):
        layer(torch.zeros(2, 4), torch.zeros(2, 4, 4), torch.ones(2), torch.ones(2))
    with pytest.raises(ValueError):
        layer(torch.zeros(2, 5), torch.zeros(2, 5, 4), torch.ones(2), torch.ones(2))
    with pytest.raises(ValueError):
        layer(torch.zeros(2, 5), torch.zeros(2, 5, 5), torch.ones(3), torch.ones(2))


def test_modules_are_picklable_and_parameter_free():
    # deepdow/callbacks.py:486 does torch.save(network): layers must hold no handles
    layers = [NumericalMarkowitz(7, max_weight=0.5), SparsemaxAllocator(7, temperature=2, max_weight=0.3),
              SoftmaxAllocator(temperature=1, formulation="variational", n_assets=7, max_weight=0.4),
              CovarianceMatrix(sqrt=False, shrinkage_strategy="identity", shrinkage_coef=0.2)]
    for layer in layers:
        clone = pickle.loads(pickle.dumps(layer))
        assert type(clone) is type(layer) and clone.__dict__.keys() == layer.__dict__.keys()
        buf = io.BytesIO()
        torch.save(layer, buf)
        assert sum(p.numel() for p in layer.parameters()) == 0
    assert issubclass(SolverError, Exception)

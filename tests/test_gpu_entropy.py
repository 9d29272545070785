"""GPU parity of the general sample_entropy kernel (probabilities, any class count) against the executed reference."""
import importlib.util
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _inputs():
    spec = importlib.util.spec_from_file_location("_mk_inputs", os.path.join(G, "class_entropy_inputs.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.class_entropy_inputs()


def test_class_entropy_golden_and_mirror():
    from contouring_uncertainty_b200 import engine as E
    from contouring_uncertainty_b200.task import sample_entropy

    p3, p1 = _inputs()
    g = np.load(os.path.join(G, "class_entropy.npz"))
    dev = torch.device("cuda:0")
    e3 = E.sample_entropy(torch.from_numpy(p3).to(dev)).cpu().numpy()
    e1 = E.sample_entropy(torch.from_numpy(p1).to(dev)).cpu().numpy()
    # fp32 tolerance: the reference averages the samples in fp32 (torch), the kernel in fp64 rounded once
    assert np.abs(e3 - g["e3"]).max() < 2e-6 and np.abs(e1 - g["e1"]).max() < 2e-6
    assert e3.dtype == np.float32 and (e3[:, :8] < 1e-6).all()
    # the reference-shaped mirror: (T_e, T_a, C, H, W) numpy in, numpy out; 0/1 masks keep the exact count-map route
    m = sample_entropy(p3[0].reshape(3, 4, 3, 64, 64))
    assert np.abs(m - g["e3"][0]).max() < 2e-6
    with pytest.raises(ValueError):
        E.sample_entropy(torch.zeros((1, 2, 9, 8, 8), device=dev))


@pytest.mark.parametrize("shape", [(1, 5, 3, 16, 16), (2, 7, 1, 8, 12), (1, 9, 2, 7, 9), (1, 6, 4, 8, 8), (1, 3, 6, 4, 4), (1, 1, 8, 5, 5)])
def test_odd_sample_counts_class_counts_and_unaligned_maps(shape):
    """Sample counts off the unroll factor, every class-count specialisation, maps whose size is not a multiple of 4."""
    from contouring_uncertainty_b200 import engine as E
    from oracle import port

    rng = np.random.default_rng(sum(shape))
    z = rng.normal(0, 2.0, shape).astype(np.float32)
    if shape[2] == 1:
        p = (1.0 / (1.0 + np.exp(-z))).astype(np.float32)
    else:
        p = np.exp(z - z.max(2, keepdims=True))
        p = (p / p.sum(2, keepdims=True)).astype(np.float32)
    got = E.sample_entropy(torch.from_numpy(p).to("cuda:0")).cpu().numpy()
    want = np.array([port.sample_entropy_general(x) for x in p])
    assert np.abs(got - want).max() < 2e-6

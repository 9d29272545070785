"""oracle/_ref (the verbatim reference copy made by oracle/make_ref.py): intact, importable under the stubs, and the port's
aggregation agrees with the reference's own loop on the reference's own samples.  Skipped where the copy is absent."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import make_ref, port, ref_runner  # noqa: E402

import synth  # noqa: E402

pytestmark = pytest.mark.skipif(not make_ref.verify(), reason="oracle/_ref absent (made by __graft_entry__.build() where /root/reference exists)")


def test_manifest_lists_only_python_sources_of_the_two_packages():
    import json
    files = json.load(open(os.path.join(make_ref.DEST, "MANIFEST.json")))["files"]
    assert len(files) > 100
    assert all(f.endswith(".py") and (f.startswith("contour_uncertainty/") or f.startswith("vital/vital/")) for f in files)


@pytest.mark.parametrize("sampler,n", [("psm", 6), ("naive", 12)])
def test_reference_unit_matches_port_on_its_own_samples(sampler, n):
    from contouring_uncertainty_b200.engine import DEFAULT_PSM
    logits = synth.heatmap_logits(1, seed=31).astype(np.float32)
    np.random.seed(5)
    _, r = ref_runner.image_unit(logits, n, sampler, DEFAULT_PSM)
    mu, cov = port.dsnt_predict_on_batch(logits, dtype=np.float32)
    assert np.abs(mu[0] - r["mu"]).max() < 2e-3 and np.allclose(cov[0], r["cov"], rtol=2e-4, atol=2e-4)   # fp32 torch sums
    o = port.uncertainty_from_samples(np.asarray(r["samples"], dtype=np.float32))
    assert np.array_equal(o["area"], r["area"])
    assert np.array_equal(o["pred"], r["pred"].reshape(256, 256))
    assert np.abs(o["mean"] - r["mean"].reshape(256, 256)).max() < 1e-15
    assert np.abs(o["entropy"] - r["entropy"].reshape(256, 256)).max() < 1e-12


def test_reference_sequence_and_skew_units_match_port_on_their_own_samples():
    """The verbatim SequencePSMSampler / SkewPosteriorShapeModelSampler units that time configs 5 / 3: shapes, and the port's
    aggregation on the reference's own samples (masks bit-exact, hence areas / pred; mean / entropy to fp64 rounding)."""
    import random

    import torch
    from contouring_uncertainty_b200.engine import DEFAULT_PSM, DEFAULT_SEQ_PSM

    mus, covs = synth.point_distributions(2, seed=3)
    ctr = mus[:1].mean(1, keepdims=True)
    mus[1] = (ctr + (mus[:1] - ctr) * 0.8)[0]
    np.random.seed(1); torch.manual_seed(1); random.seed(1)
    _, r = ref_runner.pair_unit(mus, covs, 3, DEFAULT_PSM, DEFAULT_SEQ_PSM)
    assert r["samples"].shape == (3, 2, 21, 2)
    for inst in range(2):
        o = port.uncertainty_from_samples(np.asarray(r["samples"][:, inst], dtype=np.float32))
        got = r[f"instant{inst}"]
        assert np.array_equal(o["area"], got["area"]) and np.array_equal(o["pred"], got["pred"].reshape(256, 256))
        assert np.abs(o["entropy"] - got["entropy"].reshape(256, 256)).max() < 1e-12
    logits = synth.heatmap_logits(1, seed=31).astype(np.float32)
    al = synth.skew_alphas(1, seed=32)
    _, r = ref_runner.skew_image_unit(logits, al[0], 2, DEFAULT_PSM)
    assert r["samples"].shape == (2, 21, 2)
    o = port.uncertainty_from_samples(np.asarray(r["samples"], dtype=np.float32))
    assert np.array_equal(o["area"], r["area"]) and np.abs(o["mean"] - r["mean"].reshape(256, 256)).max() < 1e-15

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

#!/usr/bin/env python
"""Time the VERBATIM reference (imported from /root/reference under tests/golden/_refstub.py) against the oracle port on the
same bounded sample of config 2, in the build container (the reference cannot travel to the GPU box: bench.py's CPU arm is
the port).  One process, one thread.  Prints one JSON line; the ratio says how conservative the port-based CPU baseline is."""
import json, os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests")); sys.path.insert(0, os.path.join(R, "tests", "golden"))
import numpy as np
import _refstub
_refstub.install()
import torch
import synth
from oracle import port
from contour_uncertainty.data.camus.utils import USContourToMask
from contour_uncertainty.sampler.posterior_shape_model.psm import PosteriorShapeModelSampler
from contour_uncertainty.task.regression.dsnt.utils import dsnt, flat_softmax, normalized_to_pixel_coordinates
from contour_uncertainty.task.uncertainty import UncertaintyTask

n_samples = int(sys.argv[1]) if len(sys.argv) > 1 else 40
psm = os.path.join(_refstub.REFERENCE_ROOT, "camus-cont_psm_11_no_std.npy")
logits = synth.heatmap_logits(1, seed=1000)
t0 = time.perf_counter()
with torch.no_grad():
    h = flat_softmax(torch.from_numpy(logits))
    coords, var, covar = dsnt(h)
    mu = normalized_to_pixel_coordinates(coords, 256)[0]
    pv, pc = var[0] * 128 ** 2, covar[0] * 128 ** 2
    cov = torch.zeros((21, 2, 2)); cov[:, 0, 0] = pv[:, 0]; cov[:, 1, 1] = pv[:, 1]; cov[:, 0, 1] = cov[:, 1, 0] = pc
    s = PosteriorShapeModelSampler(psm)(mu, cov, n=n_samples).numpy()
masks = np.array([USContourToMask.__call__(s[k], (256, 256), labels=[], apply_argmax=False) for k in range(n_samples)])
UncertaintyTask.sample_entropy(masks.astype(float))
t_ref = time.perf_counter() - t0
model = port.ShapeModel.load(psm)
z = np.random.default_rng(0).standard_normal((n_samples, 17, 2))
t0 = time.perf_counter()
m, c = port.dsnt_predict_on_batch(logits.astype(np.float32), dtype=np.float32)
sp = port.psm_sample(model, m[0], c[0], z).astype(np.float32)
port.uncertainty_from_samples(sp)
t_port = time.perf_counter() - t0
print(json.dumps({"config": 2, "samples": n_samples, "reference_samples_per_s_one_core": n_samples / t_ref,
                  "port_samples_per_s_one_core": n_samples / t_port, "port_over_reference": t_ref / t_port}))

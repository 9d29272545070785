"""Executes the VERBATIM reference (oracle/_ref, made by oracle/make_ref.py) for the CPU baseline of bench.py.

TEST / MEASUREMENT INFRASTRUCTURE ONLY: imported by bench.py's CPU legs and by tests; never by the product package.
The reference runs under the import stubs of tests/golden/_refstub.py (matplotlib, scikit-image, hydra, lightning ... are
absent from the image; `skimage.draw.line` is the restated Bresenham).  Every function below calls the
reference's own functions -- no restatement: what is timed is contour_uncertainty's code.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(HERE, "_ref")


def available() -> bool:
    """oracle/_ref holds an intact copy (sha256 manifest) of the reference."""
    from oracle import make_ref
    return make_ref.verify(REF)


_mods = None


def _load():
    """Install the stubs against oracle/_ref and import the reference modules of the dsnt-al / dsnt-al2 path."""
    global _mods
    if _mods is not None:
        return _mods
    os.environ["CU_REFERENCE_ROOT"] = REF
    for p in (os.path.join(ROOT, "tests", "golden"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import _refstub
    _refstub.REFERENCE_ROOT = REF
    _refstub.install()
    import torch
    from contour_uncertainty.data.camus.utils import USContourToMask
    from contour_uncertainty.sampler.naive import NaiveSampler
    from contour_uncertainty.sampler.posterior_shape_model.psm import PosteriorShapeModelSampler
    from contour_uncertainty.task.regression.dsnt.utils import dsnt, flat_softmax, normalized_to_pixel_coordinates
    from contour_uncertainty.task.uncertainty import UncertaintyTask
    from contour_uncertainty.utils.clinical import lv_area
    _mods = dict(torch=torch, USContourToMask=USContourToMask, NaiveSampler=NaiveSampler,
                 PosteriorShapeModelSampler=PosteriorShapeModelSampler, dsnt=dsnt, flat_softmax=flat_softmax,
                 normalized_to_pixel_coordinates=normalized_to_pixel_coordinates, UncertaintyTask=UncertaintyTask, lv_area=lv_area)
    return _mods


_sampler = {}


def image_unit(logits: np.ndarray, n_samples: int, sampler: str, psm_path: str):
    """One image of configs 1 / 2 through the reference's own code: flat_softmax + dsnt + pixel rescale
    (cu/task/regression/dsnt/utils.py:7-105, dsnt_al.py:118-131) -> NaiveSampler / PosteriorShapeModelSampler
    (cu/sampler/) -> USContourToMask per sample (cu/data/camus/utils.py:31-98; convert_to_mask's loop,
    cu/task/regression/contour_uncertainty.py:26-57) -> mean / round / sample_entropy (aleatoric.py:110-111,
    uncertainty.py:107-133) -> lv_area per sample.  logits (1, 21, 256, 256) fp32.  Returns (seconds, dict of results)."""
    m = _load()
    torch = m["torch"]
    t0 = time.perf_counter()
    with torch.no_grad():
        h = m["flat_softmax"](torch.from_numpy(logits))
        coords, var, covar = m["dsnt"](h)
        size = logits.shape[2]
        mu = m["normalized_to_pixel_coordinates"](coords, size)[0]
        pv, pc = var[0] * (size / 2) ** 2, covar[0] * (size / 2) ** 2
        cov = torch.zeros((mu.shape[0], 2, 2))
        cov[:, 0, 0] = pv[:, 0]; cov[:, 1, 1] = pv[:, 1]; cov[:, 0, 1] = cov[:, 1, 0] = pc
        if sampler == "naive":
            s = m["NaiveSampler"]()(mu.numpy(), cov.numpy(), n_samples)
        else:
            if psm_path not in _sampler:
                _sampler[psm_path] = m["PosteriorShapeModelSampler"](psm_path)
            s = _sampler[psm_path](mu, cov, n=n_samples).numpy()
    masks = np.array([m["USContourToMask"].__call__(s[k], (size, size), labels=[], apply_argmax=False) for k in range(n_samples)])
    mean = masks.mean(0)
    pred = mean.round()
    ent = m["UncertaintyTask"].sample_entropy(masks.astype(float))
    area = np.array([m["lv_area"](masks[k, 0]) for k in range(n_samples)])
    return time.perf_counter() - t0, dict(mu=mu.numpy(), cov=cov.numpy(), samples=np.asarray(s), mean=mean, pred=pred, entropy=ent, area=area)


def _maps_from_samples(m, samples, size):
    """convert_to_mask's loop + pred / sample_entropy / lv_area on (n, 21, 2) contour samples, the reference's own functions."""
    n = len(samples)
    masks = np.array([m["USContourToMask"].__call__(samples[k], (size, size), labels=[], apply_argmax=False) for k in range(n)])
    mean = masks.mean(0)
    return dict(mean=mean, pred=mean.round(), entropy=m["UncertaintyTask"].sample_entropy(masks.astype(float)),
                area=np.array([m["lv_area"](masks[k, 0]) for k in range(n)]))


def pair_unit(mu: np.ndarray, cov: np.ndarray, n_samples: int, psm_path: str, seq_psm_path: str):
    """One (ED, ES) pair of config 5 through the reference's own code: SequencePSMSampler.__call__
    (cu/sampler/posterior_shape_model/sequence_sampler.py:26-144; mu (2, 21, 2), cov (2, 21, 2, 2) fp32 -> (n, 2, 21, 2)),
    then per instant the mask loop, pred, sample_entropy and lv_area of `image_unit`.  Returns (seconds, dict)."""
    m = _load()
    torch = m["torch"]
    from contour_uncertainty.sampler.posterior_shape_model.sequence_sampler import SequencePSMSampler
    key = ("seq", psm_path, seq_psm_path)
    t0 = time.perf_counter()
    with torch.no_grad():
        if key not in _sampler:
            _sampler[key] = SequencePSMSampler(psm_path, seq_psm_path)
        s = _sampler[key](torch.from_numpy(np.asarray(mu, dtype=np.float32)), torch.from_numpy(np.asarray(cov, dtype=np.float32)), n=n_samples)
    s = np.asarray(s.detach().numpy() if hasattr(s, "detach") else s)
    out = dict(samples=s)
    for inst in range(2):
        out[f"instant{inst}"] = _maps_from_samples(m, s[:, inst], 256)
    return time.perf_counter() - t0, out


def skew_image_unit(logits: np.ndarray, alpha: np.ndarray, n_samples: int, psm_path: str):
    """One image of config 3 through the reference's own code: flat_softmax + dsnt + pixel rescale, then
    SkewPosteriorShapeModelSampler.__call__ (cu/sampler/posterior_shape_model/psm_skew.py:187-208; alpha (21, 2) as the network's
    skew head would give it) and the mask loop / maps of `image_unit`.  Returns (seconds, dict)."""
    m = _load()
    torch = m["torch"]
    from contour_uncertainty.sampler.posterior_shape_model.psm_skew import SkewPosteriorShapeModelSampler
    key = ("skew", psm_path)
    t0 = time.perf_counter()
    with torch.no_grad():
        h = m["flat_softmax"](torch.from_numpy(logits))
        coords, var, covar = m["dsnt"](h)
        size = logits.shape[2]
        mu = m["normalized_to_pixel_coordinates"](coords, size)
        pv, pc = var * (size / 2) ** 2, covar * (size / 2) ** 2
        cov = torch.zeros((1, mu.shape[1], 2, 2))
        cov[..., 0, 0] = pv[..., 0]; cov[..., 1, 1] = pv[..., 1]; cov[..., 0, 1] = cov[..., 1, 0] = pc
        if key not in _sampler:
            _sampler[key] = SkewPosteriorShapeModelSampler(psm_path)
        s = _sampler[key](mu, cov, torch.from_numpy(np.asarray(alpha, dtype=np.float32))[None], n=n_samples)
    s = np.asarray(s.detach().numpy() if hasattr(s, "detach") else s)[0]             # (n, 21, 2)
    out = _maps_from_samples(m, s, size)
    out.update(mu=mu[0].numpy(), cov=cov[0].numpy(), samples=s)
    return time.perf_counter() - t0, out


"""GPU parity of the two-structure (LV + MYO) masks against the golden made by executing the reference
(skimage.draw.polygon restated: parity unpinned for it) and the oracle."""
import os

import numpy as np
import pytest
import torch

import synth
from oracle import port

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_golden_argmax_and_stack():
    from contouring_uncertainty_b200.data.camus.utils import USContourToMask

    g = np.load(os.path.join(G, "myo.npz"))
    lm = g["landmarks"]
    labels = [0, 1, 2]                                           # Label.BG, Label.LV, Label.MYO
    got = USContourToMask.__call__(lm, (256, 256), labels=labels)
    bad = [(i, int((got[i] != g["argmax"][i]).sum())) for i in range(len(lm)) if not np.array_equal(got[i], g["argmax"][i])]
    assert not bad, bad
    stack = np.unpackbits(g["stack_linear"], axis=-1, bitorder="little")
    got = USContourToMask.__call__(lm, (256, 256), labels=labels, apply_argmax=False, reconstruction_type="linear")
    assert got.shape == stack.shape and np.array_equal(got, stack)
    one = USContourToMask.__call__(lm[2], (256, 256), labels=labels)
    assert one.shape == (256, 256) and np.array_equal(one, g["argmax"][2])


def test_random_landmarks_against_oracle():
    from contouring_uncertainty_b200 import engine as E

    lm = np.concatenate([synth.endo_epi_landmarks(40, seed=71), synth.endo_epi_landmarks(24, seed=72, wild=True)]).astype(np.float32)
    lv, myo, area = E.us_myo_masks(torch.from_numpy(lm).to("cuda:0"))
    lv, myo = E.unpack_masks(lv).cpu().numpy(), E.unpack_masks(myo).cpu().numpy()
    bad = []
    for i in range(len(lm)):
        seg = port.us_contour_to_mask_myo(lm[i], apply_argmax=False)
        if not (np.array_equal(lv[i], seg[1]) and np.array_equal(myo[i], seg[2])):
            bad.append((i, int((lv[i] != seg[1]).sum()), int((myo[i] != seg[2]).sum())))
    assert not bad, bad
    assert np.array_equal(area.cpu().numpy(), np.stack([lv.reshape(len(lm), -1).sum(1), myo.reshape(len(lm), -1).sum(1)], 1))
    assert not (lv & myo).any()


def test_myo_sample_stack_aggregation_matches_reference():
    """Labels with Label.MYO: convert_to_mask -> one-hot (N,T_e,T_a,3,H,W) stacks -> base-3 sample_entropy / argmax, against
    the reference's own ContourUncertaintyTask.convert_to_mask + sample_entropy (tests/golden/myo_task.npz; draw.polygon
    restated: parity unpinned for that function), and the bit-packed route (us_myo_masks -> mask_class_entropy)."""
    from contouring_uncertainty_b200 import engine as E
    from contouring_uncertainty_b200.data.config import DataParameters, Label
    from contouring_uncertainty_b200.task.regression.contour_uncertainty import ContourUncertaintyTask

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "myo_task.npz"))
    n, te, ta = g["contour_samples"].shape[:3]
    task = ContourUncertaintyTask(data_params=DataParameters((1, 256, 256), (42, 2), (Label.BG, Label.LV, Label.MYO)))
    pred, ps = task.convert_to_mask(g["mu"], (n, 3, 256, 256), g["contour_samples"])
    assert ps.shape == (n, te, ta, 3, 256, 256) and np.array_equal(pred, g["pred"])
    assert np.array_equal(ps.argmax(3), g["pred_samples_argmax"]) and np.array_equal(ps.sum((1, 2)), g["counts"])
    ent = np.array([task.sample_entropy(ps[i].astype(float)) for i in range(n)])
    assert np.abs(ent - g["entropy"]).max() <= 1e-12
    # the route that never materialises the stacks
    lm = torch.from_numpy(g["contour_samples"].reshape(-1, 42, 2)).cuda()
    lv, myo, _ = E.us_myo_masks(lm)
    e2, cnt, pr = E.mask_class_entropy(lv.reshape(n, te * ta, 2048), myo.reshape(n, te * ta, 2048), want_counts=True, want_pred=True)
    assert np.abs(e2.cpu().numpy() - g["entropy"]).max() <= 1e-12 and np.array_equal(cnt.cpu().numpy(), g["counts"])
    assert np.array_equal(pr.cpu().numpy(), g["counts"].argmax(1))

"""Seeded inputs of the class-entropy golden (shared by make_golden.py and the tests)."""
import numpy as np


def class_entropy_inputs(seed=81):
    """Softmax probabilities (2 images, 12 samples, 3 classes, 64 x 64) and sigmoid ones (C = 1), incl. saturated pixels."""
    rng = np.random.default_rng(seed)
    z = rng.normal(0, 3.0, (2, 12, 3, 64, 64)).astype(np.float32)
    z[:, :, 0, :8] += 40.0                                         # one class certain: entropy 0
    p3 = np.exp(z - z.max(2, keepdims=True))
    p3 = (p3 / p3.sum(2, keepdims=True)).astype(np.float32)
    p1 = (1.0 / (1.0 + np.exp(-rng.normal(0, 4.0, (2, 12, 1, 64, 64))))).astype(np.float32)
    p1[:, :, :, :4] = 0.0
    p1[:, :, :, 4:8] = 1.0
    return p3, p1

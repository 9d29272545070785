"""Mirror of contour_uncertainty/sampler/posterior_shape_model/psm_skew_sequence.py::SequenceSkewPSMSampler."""
from __future__ import annotations

from pathlib import Path
from typing import List, Optional

import torch

from ... import engine as E
from .psm_skew import SkewPosteriorShapeModelSampler


class SequenceSkewPSMSampler(SkewPosteriorShapeModelSampler):
    """psm_skew_sequence.py:21-104.  ``__call__(mu (2,K,2), cov (2,K,2,2), alpha (2,K,2), n=1) -> (2,n,K,2)``.

    Per sample a random instant is drawn first with the skew-PSM sampler; the other instant is then drawn from
    the grid pdfs ``skew pdf x N(grid; sequence-model conditional)`` (three initial points by multinomial, level
    points with the shape-model conditional on top).  Batched input (Np,2,K,2) gives (Np,2,n,K,2).
    ``first_instant`` (n,), ``z`` (n,3,3) and ``expo`` (n,31,65536) replay the reference's random draws.
    """

    def __init__(self, psm_path: Optional[Path] = None, sequence_psm_path: Optional[Path] = None, levels: int = 3,
                 skew_indices: Optional[List[int]] = None, device=None):
        super().__init__(psm_path, levels, skew_indices, device)
        self.seq = E.SequenceModelConstants(str(sequence_psm_path) if sequence_psm_path is not None else E.DEFAULT_SEQ_PSM,
                                            self.device)

    def __call__(self, mu: torch.Tensor, cov: torch.Tensor, alpha: torch.Tensor, n: int = 1, debug_img=None,
                 progress_bar: bool = False, first_instant=None, z=None, expo=None, seed: int = 0, return_indices: bool = False):
        out_dev = mu.device
        batched = mu.dim() == 4
        to = lambda t, shp: torch.as_tensor(t, dtype=torch.float32).to(self.device).reshape(shp).contiguous()
        m, c, a = to(mu, (-1, 2, E.K, 2)), to(cov, (-1, 2, E.K, 2, 2)), to(alpha, (-1, 2, E.K, 2))
        npair = m.shape[0]
        fi = None if first_instant is None else torch.as_tensor(first_instant).to(self.device, torch.uint8).reshape(npair, n).contiguous()
        zt = None if z is None else to(z, (npair, n, 3, 3))
        et = None if expo is None else to(expo, (npair, n, 31, E.HW * E.HW))
        r = E.seq_skew_sample(self.model, self.seq, m, c, a, n, first_instant=fi, z=zt, expo=et, seed=seed,
                              return_indices=return_indices)
        s = r[0] if return_indices else r
        s = (s if batched else s[0]).to(out_dev)
        return (s, r[1].to(out_dev)) if return_indices else s

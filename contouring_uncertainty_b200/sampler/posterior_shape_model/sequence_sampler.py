"""Mirror of contour_uncertainty/sampler/posterior_shape_model/sequence_sampler.py::SequencePSMSampler."""
from __future__ import annotations

from pathlib import Path
from typing import Optional

import torch

from ... import engine as E
from .psm import PosteriorShapeModelSampler


class SequencePSMSampler(PosteriorShapeModelSampler):
    """sequence_sampler.py:13-144.  ``__call__(mu (2,K,2), cov (2,K,2,2), alpha=None, n=1) -> (n,2,K,2)``.

    Per sample a random instant is drawn first (``first_instant``; the reference uses
    ``random.randint(0, 1)``), the other instant is conditioned on it through the 84-dim sequence
    PSM.  Batched input (Np,2,K,2) gives (Np,n,2,K,2).  ``z`` (n,2,17,2) replays the draws
    (z[:,0] first-sampled instant, z[:,1] the other).
    """

    def __init__(self, psm_path: Optional[Path] = None, sequence_psm_path: Optional[Path] = None, levels: int = 3, device=None):
        super().__init__(psm_path, levels, device)
        self.seq = E.SequenceModelConstants(str(sequence_psm_path) if sequence_psm_path is not None else E.DEFAULT_SEQ_PSM,
                                            self.device)

    def __call__(self, mu: torch.Tensor, cov: torch.Tensor, alpha=None, n: int = 1, debug_img=None, z=None,
                 first_instant=None, seed: int = 0) -> torch.Tensor:
        if alpha is not None:
            raise NotImplementedError("alpha (skew-normal) belongs to the dsnt-skew row")
        out_dev = mu.device
        batched = mu.dim() == 4
        m = mu.to(self.device, torch.float32).reshape(-1, 2, E.K, 2).contiguous()
        c = cov.to(self.device, torch.float32).reshape(-1, 2, E.K, 2, 2).contiguous()
        npair = m.shape[0]
        zt = None if z is None else torch.as_tensor(z, dtype=torch.float32).to(self.device).reshape(npair, n, 2, E.Z_PSM).contiguous()
        fi = None if first_instant is None else torch.as_tensor(first_instant).to(self.device, torch.uint8).reshape(npair, n).contiguous()
        s = E.sequence_sample(self.model, self.seq, m, c, n, z=zt, first_instant=fi, seed=seed)
        return (s if batched else s[0]).to(out_dev)

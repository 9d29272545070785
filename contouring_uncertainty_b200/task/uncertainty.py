"""Mirror of contour_uncertainty/task/uncertainty.py::UncertaintyTask (:24-147), the predict-side part, without Lightning.

The reference class is a ``pytorch_lightning.LightningModule`` built by Hydra from a config tree (network, optimiser, data
parameters).  Training, checkpoint IO and logging stay with the reference; this class keeps the constructor arguments and
the attributes / methods the predict path reads (``hparams.t_a / t_e / data_params``, ``model``, ``ensembling``,
``sample_entropy``, ``get_voxelspacing_and_instants``) so that ``predict_step(batch, batch_idx)`` can be called the same way.
"""
from __future__ import annotations

from pathlib import Path
from types import SimpleNamespace
from typing import List, Sequence, Union

import numpy as np
import torch

from ..data.config import CamusTags, DataParameters, Label, Tags
from . import sample_entropy as _sample_entropy


class HParams(SimpleNamespace):
    """What ``save_hyperparameters()`` leaves in ``self.hparams``: attribute access to the constructor arguments."""

    def __contains__(self, k):
        return k in self.__dict__


class UncertaintyTask:
    """uncertainty.py:24-75.  ``model``: the network (any callable mapping an image batch to the head's output), or a
    list of networks = an ensemble (what ``ensemble_ckpt`` loads in the reference, :61-76: ``t_e`` becomes their number);
    a single network with ``t_e > 1`` is called ``t_e`` times (MC dropout is the network's business: the reference patches
    its dropout layers with ``patch_module``, :80).  ``data_params``: (in_shape, out_shape, labels)."""

    def __init__(self, t_a: int = 1, t_e: int = 1, train_ensemble: bool = False, ensemble_ckpt: Union[Path, List] = None,
                 *args, model=None, data_params: DataParameters = None, device=None, **kwargs):
        if ensemble_ckpt is not None:
            raise NotImplementedError("checkpoint loading is Lightning's (out of scope): pass the loaded networks as model=[...]")
        if data_params is None:
            data_params = DataParameters(in_shape=(1, 256, 256), out_shape=(21, 2), labels=(Label.BG, Label.LV))
        self.hparams = HParams(t_a=t_a, t_e=t_e, train_ensemble=train_ensemble, ensemble_ckpt=ensemble_ckpt,
                               data_params=data_params, **kwargs)
        if isinstance(model, (list, tuple)):
            self.model = list(model)
            self.ensembling = True
            self.hparams.t_e = len(self.model)
        else:
            self.model = model
            self.ensembling = False
        self.is_val_step = False
        self.trainer = None
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())

    def forward(self, *args, **kwargs):
        return self.model(*args, **kwargs)

    __call__ = forward

    def predict_step(self, batch, batch_idx: int = 0, dataloader_idx: int = 0):
        raise NotImplementedError

    @staticmethod
    def sample_entropy(samples):
        """uncertainty.py:107-133: samples (T_e,T_a,C,H,W) or (T,C,H,W) -> (H,W) entropy of the sample mean."""
        return _sample_entropy(samples)

    @staticmethod
    def get_voxelspacing_and_instants(batch):
        """uncertainty.py:136-147."""
        if CamusTags.metadata in batch.keys():
            meta = batch[CamusTags.metadata]
            instants = meta.instants
            voxelspacing = np.array(meta.voxelspacing)
            if meta.gt is not None:
                voxelspacing = voxelspacing * meta.gt.shape / batch[Tags.gt].shape
            voxelspacing = voxelspacing[1:]
        else:
            instants = None
            voxelspacing = None
        return voxelspacing, instants

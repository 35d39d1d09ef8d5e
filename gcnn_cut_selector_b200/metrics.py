"""Ranking accuracy of the reference's training / validation loop on the device (model_trainer.py:279-302).

The reference splits predictions and targets per sample on the host, sorts both with Python's ``sorted`` and looks for
the first deviating position -- two ``.numpy()`` synchronisations and an O(n log n) Python sort per sample per batch.
Here one kernel launch handles the whole batch (``gcnn_ranking_deviation``); only ``n_samples`` integers come back.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


def ranking_deviation(predictions: torch.Tensor, improvements, n_cuts) -> torch.Tensor:
    """First position at which the predicted and the true ranking of each sample differ (int32 device tensor)."""
    lib = _lib.load()
    dev = predictions.device
    n_cuts = np.asarray(n_cuts, dtype=np.int64).reshape(-1)
    offsets = torch.from_numpy(np.concatenate([[0], np.cumsum(n_cuts)]).astype(np.int32)).to(dev)
    truth = torch.as_tensor(np.asarray(improvements, dtype=np.float32) if not torch.is_tensor(improvements) else improvements,
                            dtype=torch.float32, device=dev).contiguous()
    pred = predictions.detach().to(torch.float32).contiguous()
    if pred.numel() != int(n_cuts.sum()) or truth.numel() != pred.numel():
        raise _lib.InvalidArgumentError("predictions / improvements / n_cuts disagree on the number of cuts")
    out = torch.empty(len(n_cuts), dtype=torch.int32, device=dev)
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(lib.gcnn_ranking_deviation(pred.data_ptr(), truth.data_ptr(), offsets.data_ptr(), len(n_cuts),
                                          int(n_cuts.max()) if len(n_cuts) else 0, out.data_ptr(), st))
    return out


def ranking_accuracy(predictions: torch.Tensor, improvements, n_cuts, fractions) -> np.ndarray:
    """``acc`` of model_trainer.py:284-301: per fraction, the number of samples whose correctly ranked prefix covers at
    least that fraction of their cuts."""
    n_cuts = np.asarray(n_cuts, dtype=np.int64).reshape(-1)
    dev = ranking_deviation(predictions, improvements, n_cuts).cpu().numpy().astype(np.float64)
    frac = dev / n_cuts
    return (frac[:, None] >= np.asarray(fractions, dtype=np.float64)[None, :]).sum(axis=0).astype(np.float64)

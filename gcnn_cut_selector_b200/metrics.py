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


def select_cuts(quality, parallelism, parallelism_forced=None, p_max: float = 0.1, p_max_ub: float = 0.5,
                max_selected: int | None = None):
    """The ranking + parallelism filter of the reference's SCIP plug-in on the device
    (``CustomCutsel.cutselselect``, model_benchmarker.py:112-157): ``quality`` [n] are the predicted bound improvements,
    ``parallelism`` [n, n] / ``parallelism_forced`` [n_forced, n] what ``getRowParallelism`` returns for every pair.
    Returns ``(order, n_selected)``: ``order`` (int32 device tensor) is the cut index at each position of the reference's
    ``sorted_cuts`` and ``n_selected`` (int32 device tensor [1]) its ``nselectedcuts``."""
    lib = _lib.load()
    q = quality if torch.is_tensor(quality) else torch.as_tensor(np.asarray(quality, dtype=np.float32))
    if not q.is_cuda:
        q = q.cuda()
    dev = q.device
    q = q.detach().to(torch.float32).contiguous()
    n = q.numel()
    as_dev = lambda x: torch.as_tensor(np.asarray(x, dtype=np.float32) if not torch.is_tensor(x) else x,
                                       dtype=torch.float32, device=dev).contiguous()
    par = as_dev(parallelism)
    if par.numel() != n * n:
        raise _lib.InvalidArgumentError("parallelism must be [n_cuts, n_cuts]")
    n_forced, pf = 0, None
    if parallelism_forced is not None and len(parallelism_forced) > 0:
        pf = as_dev(parallelism_forced)
        if pf.numel() % max(n, 1) != 0:
            raise _lib.InvalidArgumentError("parallelism_forced must be [n_forced, n_cuts]")
        n_forced = pf.numel() // max(n, 1)
    order = torch.empty(n, dtype=torch.int32, device=dev)
    n_sel = torch.empty(1, dtype=torch.int32, device=dev)
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    with torch.cuda.device(dev):
        _lib.check(lib.gcnn_select_cuts(q.data_ptr(), pf.data_ptr() if pf is not None else None, par.data_ptr(), n,
                                        n_forced, float(p_max), float(p_max_ub),
                                        int(max_selected) if max_selected is not None else n, order.data_ptr(),
                                        n_sel.data_ptr(), st))
    return order, n_sel

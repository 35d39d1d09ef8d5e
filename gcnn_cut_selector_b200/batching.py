"""Batching contract of the reference's ``utils.load_batch`` (utils.py:339-426), host side.

Samples are gzip-pickled ``{'data': [state, improvements]}`` files (data_collector.py:135-140).  Graphs are
concatenated block-diagonally: every sample's ``[2, E]`` edge-index block is shifted by the number of constraints /
cuts (row 0) and variables (row 1) that precede it (utils.py:403-407); features are cast to fp32 and indices to int32
(utils.py:413-423).  Integer arithmetic is exact (int64 shifts, checked int32 narrowing).
"""
from __future__ import annotations

import gzip
import pickle

import numpy as np

_INT32_MAX = np.iinfo(np.int32).max


def _exclusive_cumsum(counts) -> np.ndarray:
    out = np.zeros(len(counts), dtype=np.int64)
    if len(counts) > 1:
        np.cumsum(np.asarray(counts[:-1], dtype=np.int64), out=out[1:])
    return out


def _offset_edges(blocks, left_counts, var_counts) -> np.ndarray:
    left_off, var_off = _exclusive_cumsum(left_counts), _exclusive_cumsum(var_counts)
    total = sum(b.shape[1] for b in blocks)
    out = np.empty((2, total), dtype=np.int64)
    pos = 0
    for j, b in enumerate(blocks):
        e = b.shape[1]
        out[0, pos:pos + e] = np.asarray(b[0], dtype=np.int64) + left_off[j]
        out[1, pos:pos + e] = np.asarray(b[1], dtype=np.int64) + var_off[j]
        pos += e
    if total and (out.max() > _INT32_MAX or out.min() < -_INT32_MAX - 1):
        raise OverflowError("edge index does not fit int32 (utils.py:414 casts to tf.int32)")
    return out.astype(np.int32)


def concat_samples(samples):
    """Stack in-memory samples ``(state, improvements)`` into the 11-tuple ``load_batch`` returns:
    (cons_feats, cons_edge_inds, cons_edge_feats, var_feats, cut_feats, cut_edge_inds, cut_edge_feats,
     n_cons, n_vars, n_cuts, improvements)."""
    states = [s for s, _ in samples]
    n_cons = [s[0]["values"].shape[0] for s in states]
    n_vars = [s[2]["values"].shape[0] for s in states]
    n_cuts = [s[3]["values"].shape[0] for s in states]

    def stack(i):
        return np.concatenate([s[i]["values"] for s in states], axis=0).astype(np.float32)

    cons_ei = _offset_edges([s[1]["indices"] for s in states], n_cons, n_vars)
    cut_ei = _offset_edges([s[4]["indices"] for s in states], n_cuts, n_vars)
    improvements = np.concatenate([np.asarray(imp) for _, imp in samples]).astype(np.float32)
    return (stack(0), cons_ei, stack(1), stack(2), stack(3), cut_ei, stack(4),
            np.asarray(n_cons, dtype=np.int32), np.asarray(n_vars, dtype=np.int32),
            np.asarray(n_cuts, dtype=np.int32), improvements)


def load_batch(sample_files):
    """Drop-in for ``utils.load_batch`` (returns numpy arrays instead of TF tensors)."""
    samples = []
    for filename in sample_files:
        if isinstance(filename, bytes):  # tf.data hands file names over as bytes (utils.py:334)
            filename = filename.decode()
        with gzip.open(filename, "rb") as fh:
            state, improvements = pickle.load(fh)["data"]
        samples.append((state, improvements))
    return concat_samples(samples)


def model_inputs(batch, per_sample_counts: bool = False):
    """The 10-tuple the model takes: features/indices plus the three *totals* (model_trainer.py:259-263).  With
    ``per_sample_counts`` the last three entries stay the per-sample vectors ``load_batch`` returns; the B200 model
    accepts either and uses the vectors to stage per-sample tables in shared memory."""
    if per_sample_counts:
        return tuple(batch[:10])
    return tuple(batch[:7]) + (int(np.sum(batch[7])), int(np.sum(batch[8])), int(np.sum(batch[9])))

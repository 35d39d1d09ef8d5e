"""CPU (no GPU needed): the C-ABI library loads and exports every symbol the header declares; host-side batching
and synthetic generators honour the reference's contracts."""
import os
import pickle
import re

import numpy as np
import pytest

import gcnn_oracle as orc
from gcnn_cut_selector_b200 import _lib, batching, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from gcnn_cut_selector_b200 import build
    build.build()
    return _lib.load()


def test_header_symbols_all_exported_and_bound(lib):
    header = open(os.path.join(ROOT, "include", "gcnn_b200.h")).read()
    declared = set(re.findall(r"\b(gcnn_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)


def test_version_and_param_table_match_reference_order(lib):
    assert lib.gcnn_version() >= 100
    table = _lib.param_table()
    assert [(n, tuple(s), t) for n, s, t, _ in table] == [(n, tuple(s), t) for n, s, t in orc.PARAM_SPECS]
    # offsets are the running sums inside each flat buffer
    o_t = o_p = 0
    for name, shape, trainable, off in table:
        k = int(np.prod(shape))
        if trainable:
            assert off == o_t, name
            o_t += k
        else:
            assert off == o_p, name
            o_p += k
    assert (o_t, o_p) == (_lib.N_TRAINABLE, _lib.N_PRENORM) == (93121, 58)


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libgcnn_b200.so")
    with pytest.raises(ImportError, match="no CPU fallback"):
        _lib.load()


def test_invalid_param_index_reports_error(lib):
    with pytest.raises(_lib.InvalidArgumentError):
        _lib.check(lib.gcnn_param_info(99, None, 0, None, None, None, None))


def test_batching_matches_reference_load_batch(golden_dir, tmp_path):
    with open(os.path.join(golden_dir, "batch_tiny_samples.pkl"), "rb") as fh:
        samples = pickle.load(fh)
    want = np.load(os.path.join(golden_dir, "batch_tiny.npz"))
    got = batching.concat_samples(samples)
    for i, g in enumerate(got):
        w = want[f"out{i}"]
        assert g.dtype == w.dtype and g.shape == w.shape, i
        np.testing.assert_array_equal(g, w)
    import gzip
    files = []
    for i, (state, imp) in enumerate(samples):
        path = str(tmp_path / f"sample_{i}.pkl")
        with gzip.open(path, "wb") as fh:
            pickle.dump({"data": [state, imp]}, fh)
        files.append(path.encode() if i == 0 else path)  # tf.data passes bytes
    for i, g in enumerate(batching.load_batch(files)):
        np.testing.assert_array_equal(g, want[f"out{i}"])


def test_batching_single_and_ragged():
    samples = synth.make_samples("tiny", 1, seed0=3)
    one = batching.concat_samples(samples)
    np.testing.assert_array_equal(one[1], samples[0][0][1]["indices"].astype(np.int32))
    # an edge-less sample in the middle keeps the offsets of the following one right
    mid = synth.make_samples("tiny", 3, seed0=9)
    (c, ce, v, k, ke), imp = mid[1]
    empty = {"indices": np.zeros((2, 0), np.int64), "values": np.zeros((0, 1))}
    mid[1] = ((c, empty, v, k, empty), imp)
    got = batching.concat_samples(mid)
    ref = orc.concat_samples(mid)
    for g, r in zip(got, ref):
        np.testing.assert_array_equal(g, r)


def test_batching_int32_overflow_is_an_error():
    with pytest.raises(OverflowError):
        batching._offset_edges([np.array([[0], [0]]), np.array([[5], [5]])], [2 ** 31, 1], [1, 1])


@pytest.mark.parametrize("shape,n_edges", [("setcov", 25000), ("capfac", 40200), ("combauc", 2800), ("indset", 3916)])
def test_synth_shapes(shape, n_edges):
    (cons, ce, var, cut, ke), imp = synth.make_sample(shape, seed=1)
    n_cons, n_vars, n_cuts, nnz = synth.SHAPES[shape]
    assert cons["values"].shape == (n_cons, 4) and var["values"].shape == (n_vars, 14)
    assert cut["values"].shape == (n_cuts, 6) and imp.shape == (n_cuts,)
    ei = ce["indices"]
    assert ei.shape == (2, n_edges) and ce["values"].shape == (n_edges, 1)
    assert np.all(np.diff(ei[0]) >= 0), "row-major sorted like get_state (utils.py:102-104)"
    assert ei[0].max() < n_cons and ei[1].max() < n_vars and ei.min() >= 0
    pairs = ei[0].astype(np.int64) * n_vars + ei[1]
    assert np.unique(pairs).size == n_edges
    assert ke["indices"].shape == (2, n_cuts * nnz)


def test_batch_struct_matches_the_library(lib):
    """The binding's gcnn_batch (ctypes) and the library's agree in size; field order is the header's."""
    import ctypes as C
    assert lib.gcnn_batch_bytes() == C.sizeof(_lib.Batch)
    hdr = open(os.path.join(ROOT, "include", "gcnn_b200.h")).read()
    body = hdr[hdr.index("typedef struct gcnn_batch {"):hdr.index("} gcnn_batch;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = [n for stmt in body.split(";") for n in re.findall(r"[\*\s,](\w+)\s*(?=,|$)", stmt.strip())]
    assert names == [f for f, _ in _lib.Batch._fields_]


def test_host_batch_packing_and_local_columns(monkeypatch):
    """HostBatch's host-side logic without a GPU (pin_memory stubbed): every packed section at a 256-byte offset inside
    the one buffer and equal to its source array; row pointers index the sorted list; uint16 local columns + the sample's
    first variable reproduce the column indices; an unsorted list and a batch without per-sample counts keep full indices;
    the byte count is what the library would copy."""
    import torch
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)
    from gcnn_cut_selector_b200 import HostBatch
    batch = list(batching.concat_samples(synth.make_samples("setcov", 3, seed0=1)))
    perm = np.random.default_rng(0).permutation(batch[5].shape[1])
    shuffled = list(batch)
    shuffled[5], shuffled[6] = batch[5][:, perm], batch[6][perm]
    h = HostBatch(tuple(batch), row_pointers=True)
    base, size = h.arena.data_ptr(), h.arena.numel()
    inside = [x for x in h.tensors + [h.targets] + h.row_ptrs + h.col16 if base <= x.data_ptr() < base + size]
    assert len(inside) == 10 and all((x.data_ptr() - base) % 256 == 0 for x in inside)
    assert h.batch.packed == base and h.batch.packed_bytes == h.h2d_bytes <= size
    for k in range(7):
        np.testing.assert_array_equal(h.tensors[k].numpy().reshape(-1), np.asarray(batch[k]).reshape(-1))
    np.testing.assert_array_equal(h.targets.numpy(), batch[10])
    var_off = np.concatenate(([0], np.cumsum(batch[8])))
    for i, (k, left_counts) in enumerate(((1, batch[7]), (5, batch[9]))):
        rows, cols = batch[k][0], batch[k][1]
        rp = h.row_ptrs[i].numpy()
        assert rp[0] == 0 and rp[-1] == rows.size
        np.testing.assert_array_equal(np.repeat(np.arange(rp.size - 1), np.diff(rp)), rows)
        sample = np.searchsorted(np.cumsum(left_counts), rows, side="right")
        np.testing.assert_array_equal(h.col16[i].numpy().astype(np.int64) + var_off[sample], cols)
    plain = HostBatch(tuple(batch), row_pointers=False, packed=False)
    assert plain.arena is None and plain.batch.packed is None and plain.row_ptrs == [None, None]
    assert h.h2d_bytes < plain.h2d_bytes - 5 * (batch[1].shape[1] + batch[5].shape[1])
    hs = HostBatch(tuple(shuffled), row_pointers=True)
    assert hs.row_ptrs[0] is not None and hs.row_ptrs[1] is None and hs.col16[1] is None
    totals = list(batch)
    totals[7:10] = [int(np.sum(x)) for x in totals[7:10]]
    ht = HostBatch(tuple(totals), row_pointers=True)
    assert ht.row_ptrs[0] is not None and ht.col16 == [None, None] and ht.batch.n_samples == 0
    assert HostBatch(tuple(batch)).row_ptrs == [None, None]  # default: only lists of >= 128 k edges


def test_integration_stub_declares_the_same_batch_struct():
    """INTEGRATION.md's ctypes stub (what a maintainer of the reference would paste) lists gcnn_batch's fields exactly as the
    binding does -- a shorter struct would make the library read past the caller's memory."""
    txt = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blk = txt[txt.index("class Batch(C.Structure):"):txt.index("lib.gcnn_batch_bytes.restype")]
    assert re.findall(r'\("(\w+)", C\.', blk) == [f for f, _ in _lib.Batch._fields_]

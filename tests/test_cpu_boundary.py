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

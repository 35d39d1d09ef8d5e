"""Packed sample records (gcnn_cut_selector_b200/shards.py, csrc/records.cu): the record format round-trips the reference's
samples, and the device-side batch assembly is bit-exact with the reference's own ``utils.load_batch`` output
(tests/golden/batch_tiny.npz, written by oracle/make_golden.py from the unmodified reference) and with the host-side
restatement on larger synthetic batches."""
import ctypes as C
import gzip
import os
import pickle

import numpy as np
import pytest
import torch

import gcnn_oracle as orc
from gcnn_cut_selector_b200 import _lib, batching, shards, synth


def _golden_samples(golden_dir):
    with open(os.path.join(golden_dir, "batch_tiny_samples.pkl"), "rb") as fh:
        return pickle.load(fh)


def _edgeless(sample):
    (c, _, v, k, _), imp = sample
    empty = {"indices": np.zeros((2, 0), np.int64), "values": np.zeros((0, 1))}
    return (c, empty, v, k, empty), imp


# ---- CPU: format ------------------------------------------------------------------------------------------------------
def test_record_round_trip_reproduces_reference_batch(golden_dir):
    samples = _golden_samples(golden_dir)
    want = np.load(os.path.join(golden_dir, "batch_tiny.npz"))
    unpacked = [shards.unpack_record(shards.pack_sample(s, imp)) for s, imp in samples]
    got = batching.concat_samples(unpacked)
    for i, g in enumerate(got):
        assert g.dtype == want[f"out{i}"].dtype
        np.testing.assert_array_equal(g, want[f"out{i}"])


@pytest.mark.parametrize("compress", [True, False])
def test_record_flags_sizes_and_unsorted_lists(compress):
    from gcnn_cut_selector_b200 import build
    build.build()
    base = synth.make_samples("mini", 2, seed0=11)
    cases = [base[0], synth.shuffle_edges(base[1], 5), _edgeless(base[0])]
    for k, (state, imp) in enumerate(cases):
        rec = shards.pack_sample(state, imp, compress_rows=compress)
        flags, nc, nv, nk, ec, ek, total = shards.record_counts(rec)
        assert total == len(rec) and len(rec) % 16 == 0
        assert total == shards.record_bytes(nc, nv, nk, ec, ek, flags)  # the library's layout agrees with the packer's
        sorted_bits = shards.CONS_ROWS_SORTED | shards.CUT_ROWS_SORTED
        if k == 1:
            assert not flags & (shards.CONS_ROWS_AS_PTR | shards.CUT_ROWS_AS_PTR)  # shuffled lists keep their row indices
        else:
            assert flags & sorted_bits == sorted_bits
            if compress and k == 0:
                assert flags & shards.CONS_ROWS_AS_PTR
        (c, ce, v, kk, ke), imp2 = shards.unpack_record(rec)
        np.testing.assert_array_equal(ce["indices"], np.asarray(state[1]["indices"], dtype=np.int32))
        np.testing.assert_array_equal(ke["indices"], np.asarray(state[4]["indices"], dtype=np.int32))
        np.testing.assert_array_equal(c["values"], np.asarray(state[0]["values"], dtype=np.float32))
        np.testing.assert_array_equal(ce["values"], np.asarray(state[1]["values"], dtype=np.float32))
        np.testing.assert_array_equal(imp2, np.asarray(imp, dtype=np.float32))


def test_out_of_range_local_index_is_rejected():
    (c, ce, v, k, ke), imp = synth.make_samples("tiny", 1, seed0=2)[0]
    bad = {"indices": ce["indices"].copy(), "values": ce["values"]}
    bad["indices"][1, 0] = v["values"].shape[0]
    with pytest.raises(ValueError):
        shards.pack_sample((c, bad, v, k, ke), imp)


def test_shard_file_and_conversion_from_reference_sample_files(golden_dir, tmp_path):
    samples = _golden_samples(golden_dir)
    files = []
    for i, (state, imp) in enumerate(samples):
        path = str(tmp_path / f"sample_{i}.pkl")
        with gzip.open(path, "wb") as fh:
            pickle.dump({"data": [state, imp]}, fh)  # data_collector.py:135-140
        files.append(path)
    shard = str(tmp_path / "train.shard")
    assert shards.convert_sample_files(files, shard) == len(samples)
    reader = shards.ShardReader(shard, pin=False)
    assert len(reader) == len(samples)
    want = np.load(os.path.join(golden_dir, "batch_tiny.npz"))
    got = batching.concat_samples([reader.sample(i) for i in range(len(reader))])
    for i, g in enumerate(got):
        np.testing.assert_array_equal(g, want[f"out{i}"])
    ids = list(range(len(reader)))
    assert reader.totals(ids) == (want["out0"].shape[0], want["out3"].shape[0], want["out4"].shape[0],
                                  want["out1"].shape[1], want["out5"].shape[1])
    ptrs = reader.pointers(ids)
    assert np.all(np.diff(ptrs.astype(np.int64)) == np.diff(reader.offsets)[:-1])  # neighbours in memory


def _random_sample(rng, n_cons, n_vars, n_cuts, ec, ek, sort_rows):
    def edges(n_left, e):
        ei = np.vstack([rng.integers(0, max(n_left, 1), e), rng.integers(0, max(n_vars, 1), e)]).astype(np.int64)
        if sort_rows:
            ei = ei[:, np.argsort(ei[0], kind="stable")]
        return {"indices": ei, "values": rng.standard_normal((e, 1))}
    ec = ec if n_cons and n_vars else 0
    ek = ek if n_cuts and n_vars else 0
    return ({"values": rng.standard_normal((n_cons, 4))}, edges(n_cons, ec), {"values": rng.standard_normal((n_vars, 14))},
            {"values": rng.standard_normal((n_cuts, 6))}, edges(n_cuts, ek)), rng.uniform(0, 0.1, n_cuts)


def test_records_round_trip_property():
    """Random ragged samples (empty node sets, empty edge lists, sorted and unsorted rows, rows stored as indices or as a
    pointer): pack -> unpack is the identity on load_batch's element types, and batches of unpacked records equal the
    oracle's restatement of utils.py:395-423 on the original samples."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(st.integers(0, 2 ** 32 - 1), st.lists(st.tuples(st.integers(0, 9), st.integers(0, 12), st.integers(0, 5),
                                                         st.integers(0, 40), st.integers(0, 25), st.booleans(),
                                                         st.booleans()), min_size=1, max_size=4))
    def check(seed, specs):
        rng = np.random.default_rng(seed)
        samples, unpacked = [], []
        for n_cons, n_vars, n_cuts, ec, ek, sort_rows, compress in specs:
            sample = _random_sample(rng, n_cons, n_vars, n_cuts, ec, ek, sort_rows)
            rec = shards.pack_sample(*sample, compress_rows=compress)
            flags, nc, nv, nk, e1, e2, total = shards.record_counts(rec)
            assert (nc, nv, nk) == (n_cons, n_vars, n_cuts) and total == len(rec)
            if flags & shards.CONS_ROWS_AS_PTR:
                assert compress and flags & shards.CONS_ROWS_SORTED
            samples.append(sample)
            unpacked.append(shards.unpack_record(rec))
        want, got = orc.concat_samples(samples), batching.concat_samples(unpacked)
        for w, g in zip(want, got):
            np.testing.assert_array_equal(np.asarray(w), g)

    check()


# ---- GPU: assembly on the device --------------------------------------------------------------------------------------
def _staged_tensors(model, slot):
    from gcnn_cut_selector_b200.model import Batch
    b, tgt = Batch(), C.c_void_p()
    _lib.check(model._lib.gcnn_staged_batch(model._ws, slot, C.byref(b), C.byref(tgt), model._stream()))
    torch.cuda.synchronize()

    class _DevArray:  # a raw device pointer as a CUDA array
        def __init__(self, ptr, n, typestr):
            self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (int(ptr), False), "version": 2}

    def dev(ptr, n, dtype):
        if n == 0:
            return np.zeros(0, dtype=np.float32 if dtype == torch.float32 else np.int32)
        return torch.as_tensor(_DevArray(ptr, n, "<f4" if dtype == torch.float32 else "<i4"), device=model.device).cpu().numpy()

    nc, nv, nk, ec, ek = b.n_cons, b.n_vars, b.n_cuts, b.n_cons_edges, b.n_cut_edges
    return (dev(b.cons_feats, nc * 4, torch.float32).reshape(nc, 4), dev(b.cons_edge_inds, 2 * ec, torch.int32).reshape(2, ec),
            dev(b.cons_edge_feats, ec, torch.float32).reshape(ec, 1), dev(b.var_feats, nv * 14, torch.float32).reshape(nv, 14),
            dev(b.cut_feats, nk * 6, torch.float32).reshape(nk, 6), dev(b.cut_edge_inds, 2 * ek, torch.int32).reshape(2, ek),
            dev(b.cut_edge_feats, ek, torch.float32).reshape(ek, 1), dev(tgt.value, nk, torch.float32)), int(b.flags)


def _write_shard(tmp_path, samples, name="s.shard", **kw):
    path = str(tmp_path / name)
    shards.write_shard(path, samples, **kw)
    return shards.ShardReader(path)


@pytest.mark.gpu
def test_device_assembly_matches_reference_load_batch(golden_dir, tmp_path):
    from gcnn_cut_selector_b200 import GCNN
    model = GCNN(device=torch.device("cuda:0"), seed=0)
    samples = _golden_samples(golden_dir)
    want = np.load(os.path.join(golden_dir, "batch_tiny.npz"))
    reader = _write_shard(tmp_path, samples)
    model.stage_records(reader, list(range(len(reader))), slot=0)
    got, flags = _staged_tensors(model, 0)
    for i in range(7):
        assert got[i].dtype == want[f"out{i}"].dtype
        np.testing.assert_array_equal(got[i], want[f"out{i}"])
    np.testing.assert_array_equal(got[7], want["out10"])
    assert flags == 0  # the fixture's second sample is not sorted by row, so the batch carries no sortedness promise


@pytest.mark.gpu
@pytest.mark.parametrize("resident", [False, True], ids=["host_shard", "resident_shard"])
@pytest.mark.parametrize("compress", [True, False])
def test_device_assembly_ragged_shuffled_and_subsets(tmp_path, compress, resident):
    """Batches assembled on the device from packed records equal utils.load_batch's concatenation bit for bit; with the
    shard resident in device memory (``to_device``) the same kernel reads the records in place and only descriptors
    cross PCIe."""
    from gcnn_cut_selector_b200 import GCNN
    model = GCNN(device=torch.device("cuda:0"), seed=0)
    base = synth.make_samples("setcov", 3, seed0=21) + synth.make_samples("combauc", 2, seed0=4) + \
        synth.make_samples("skewed", 1, seed0=8)
    base[1] = synth.shuffle_edges(base[1], 3)
    base[3] = _edgeless(base[3])
    reader = _write_shard(tmp_path, base, compress_rows=compress)
    if resident:
        reader.to_device("cuda:0")
    for slot, ids in ((0, [0, 1, 2, 3, 4, 5]), (1, [4, 0, 5]), (0, [2]), (1, [3, 3])):
        staged = model.stage_records(reader, ids, slot=slot)
        got, flags = _staged_tensors(model, slot)
        want = orc.concat_samples([reader.sample(i) for i in ids])  # line-by-line restatement of utils.py:395-423
        host = batching.concat_samples([base[i] for i in ids])
        for i in range(7):
            np.testing.assert_array_equal(got[i], np.asarray(want[i]))
            np.testing.assert_array_equal(got[i], host[i])
        np.testing.assert_array_equal(got[7], host[10])
        assert flags == (0 if 1 in ids else 3)
        if resident:
            assert 0 < staged.h2d_bytes <= 256 * len(ids)  # descriptors only
        else:
            assert staged.h2d_bytes >= reader.record_bytes(ids)
    if compress:  # the row pointer takes a third off the edge bytes of sorted lists
        plain = _write_shard(tmp_path, base, name="plain.shard", compress_rows=False)
        assert reader.record_bytes([0]) < 0.75 * plain.record_bytes([0])


@pytest.mark.gpu
def test_steps_from_records_equal_steps_from_host_batches(tmp_path):
    from gcnn_cut_selector_b200 import GCNN, HostBatch
    samples = synth.make_samples("setcov", 4, seed0=77)
    reader = _write_shard(tmp_path, samples)
    losses = []
    for use_records in (False, True, "resident"):
        if use_records == "resident":
            reader.to_device("cuda:0")
        model = GCNN(device=torch.device("cuda:0"), seed=5)
        out = []
        for step, ids in enumerate(([0, 1], [2, 3], [1, 3])):
            slot = step & 1
            if use_records:
                model.stage_records(reader, ids, slot=slot)
            else:
                model.stage_host(HostBatch(batching.concat_samples([samples[i] for i in ids])), slot)
            out.append(model.train_step_staged(slot, 1e-3))
        if use_records:
            model.stage_records(reader, [0, 2], slot=1, training=False)
        else:
            model.stage_host(HostBatch(batching.concat_samples([samples[0], samples[2]])), 1, training=False)
        out.append(model.score_staged(1).copy())
        losses.append(out)
    for other in losses[1:]:
        assert losses[0][:3] == other[:3]  # same tensors in, same kernels: bit-identical losses
        np.testing.assert_array_equal(losses[0][3], other[3])


@pytest.mark.gpu
def test_bad_record_is_reported(tmp_path):
    from gcnn_cut_selector_b200 import GCNN
    model = GCNN(device=torch.device("cuda:0"), seed=0)
    reader = _write_shard(tmp_path, synth.make_samples("tiny", 2, seed0=1))
    model.stage_records(reader, [0, 1], slot=0)
    reader._view[int(reader.offsets[1])] ^= 0xFF  # corrupt the second record's magic
    with pytest.raises(_lib.InvalidArgumentError, match="bad magic"):
        model.stage_records(reader, [0, 1], slot=0)

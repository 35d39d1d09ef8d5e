"""Packed sample records and shard files: the on-disk / in-memory sample format of the B200 path.

The reference stores one gzip-pickled ``{'data': [state, improvements]}`` file per sample (data_collector.py:135-140)
and rebuilds every batch on the host: un-gzip, un-pickle, ``np.concatenate``, int64 index shifts, casts
(utils.load_batch, utils.py:339-426).  Here a sample is packed ONCE into a binary *record* whose arrays already have the
element types ``load_batch`` casts to (fp32 features, int32 indices local to the sample; layout in
``include/gcnn_b200.h``); a *shard* is a file of records with an offset table.  A batch is then a list of record
pointers: ``GCNN.stage_records`` copies the records to the device as they are and one kernel concatenates, shifts and
narrows (``csrc/records.cu``) -- bit-exact with ``load_batch``.

Edge lists sorted by row -- every list the reference produces (csr -> coo order, utils.py:102-104) -- store a row pointer
of n + 1 entries instead of E row indices, which takes a third off the bytes per edge.

Host-side numpy only.  ``unpack_record`` inverts ``pack_sample`` (tests, and conversion back to the reference format).
"""
from __future__ import annotations

import gzip
import pickle
import struct

import numpy as np

from . import _lib

MAGIC = 0x31524347           # "GCR1"
SHARD_MAGIC = 0x31534347     # "GCS1"
HEADER_BYTES = 64
CONS_ROWS_AS_PTR, CUT_ROWS_AS_PTR, CONS_ROWS_SORTED, CUT_ROWS_SORTED = 1, 2, 4, 8
CONS_F, VAR_F, CUT_F = 4, 14, 6
_INT32_MAX = np.iinfo(np.int32).max


def _pad16(b: bytes) -> bytes:
    return b + b"\0" * (-len(b) % 16)


def _local_index(a, n, what) -> np.ndarray:
    a = np.asarray(a)
    if a.size and (a.min() < 0 or a.max() >= max(n, 1) or a.max() > _INT32_MAX):
        raise ValueError(f"{what} index outside [0, {n})")
    return a.astype(np.int32)


def _rows_section(rows: np.ndarray, n_rows: int, compress: bool):
    """(section bytes, as_ptr, sorted): a row pointer for lists sorted by row, the row indices otherwise."""
    is_sorted = bool(rows.size <= 1 or np.all(rows[1:] >= rows[:-1]))
    if is_sorted and compress and rows.size > n_rows + 1:
        ptr = np.zeros(n_rows + 1, dtype=np.int64)
        np.cumsum(np.bincount(rows, minlength=n_rows), out=ptr[1:])
        return ptr.astype(np.int32).tobytes(), True, True
    return rows.tobytes(), False, is_sorted


def pack_sample(state, improvements, compress_rows: bool = True) -> bytes:
    """One sample ``(state, improvements)`` of the reference format (utils.py:236-238) as a record."""
    cons, cons_e, var, cut, cut_e = state
    f32 = lambda a, cols: np.ascontiguousarray(np.asarray(a, dtype=np.float32).reshape(-1, cols))
    cons_f, var_f, cut_f = f32(cons["values"], CONS_F), f32(var["values"], VAR_F), f32(cut["values"], CUT_F)
    n_cons, n_vars, n_cuts = cons_f.shape[0], var_f.shape[0], cut_f.shape[0]
    imp = np.asarray(improvements, dtype=np.float32).reshape(-1)
    if imp.shape[0] != n_cuts:
        raise ValueError("one improvement per cut expected")
    cei, kei = np.asarray(cons_e["indices"]), np.asarray(cut_e["indices"])
    cef, kef = f32(cons_e["values"], 1), f32(cut_e["values"], 1)
    ec, ek = cei.shape[1], kei.shape[1]
    if cef.shape[0] != ec or kef.shape[0] != ek:
        raise ValueError("one feature per edge expected")
    c_rows, c_cols = _local_index(cei[0], n_cons, "constraint"), _local_index(cei[1], n_vars, "variable")
    k_rows, k_cols = _local_index(kei[0], n_cuts, "cut"), _local_index(kei[1], n_vars, "variable")
    c_sec, c_ptr, c_sorted = _rows_section(c_rows, n_cons, compress_rows)
    k_sec, k_ptr, k_sorted = _rows_section(k_rows, n_cuts, compress_rows)
    flags = (CONS_ROWS_AS_PTR * c_ptr) | (CUT_ROWS_AS_PTR * k_ptr) | (CONS_ROWS_SORTED * c_sorted) | (CUT_ROWS_SORTED * k_sorted)
    body = b"".join(_pad16(x) for x in (cons_f.tobytes(), var_f.tobytes(), cut_f.tobytes(), imp.tobytes(), cef.tobytes(),
                                        kef.tobytes(), c_sec, c_cols.tobytes(), k_sec, k_cols.tobytes()))
    total = HEADER_BYTES + len(body)
    header = struct.pack("<8iq24x", MAGIC, flags, n_cons, n_vars, n_cuts, ec, ek, 0, total)
    assert len(header) == HEADER_BYTES
    return header + body


def record_counts(buf, offset: int = 0):
    """(flags, n_cons, n_vars, n_cuts, n_cons_edges, n_cut_edges, record_bytes) of the record at ``buf[offset:]``."""
    magic, flags, nc, nv, nk, ec, ek, _, total = struct.unpack_from("<8iq", buf, offset)
    if magic != MAGIC:
        raise ValueError("not a sample record")
    return flags, nc, nv, nk, ec, ek, total


def unpack_record(buf, offset: int = 0):
    """Inverse of ``pack_sample``: ``(state, improvements)`` with fp32 / int32 arrays."""
    flags, nc, nv, nk, ec, ek, _ = record_counts(buf, offset)
    pos = offset + HEADER_BYTES

    def take(dtype, count):
        nonlocal pos
        a = np.frombuffer(buf, dtype=dtype, count=count, offset=pos).copy()
        pos += (a.nbytes + 15) & ~15
        return a

    cons, var, cut = take(np.float32, nc * CONS_F).reshape(nc, CONS_F), take(np.float32, nv * VAR_F).reshape(nv, VAR_F), \
        take(np.float32, nk * CUT_F).reshape(nk, CUT_F)
    imp, cef, kef = take(np.float32, nk), take(np.float32, ec).reshape(ec, 1), take(np.float32, ek).reshape(ek, 1)

    def rows(as_ptr, n_rows, n_edges):
        if as_ptr:
            ptr = take(np.int32, n_rows + 1)
            return np.repeat(np.arange(n_rows, dtype=np.int32), np.diff(ptr))
        return take(np.int32, n_edges)

    c_rows = rows(flags & CONS_ROWS_AS_PTR, nc, ec)
    c_cols = take(np.int32, ec)
    k_rows = rows(flags & CUT_ROWS_AS_PTR, nk, ek)
    k_cols = take(np.int32, ek)
    state = ({"values": cons}, {"indices": np.vstack([c_rows, c_cols]), "values": cef}, {"values": var},
             {"values": cut}, {"indices": np.vstack([k_rows, k_cols]), "values": kef})
    return state, imp


def write_shard(path: str, samples, compress_rows: bool = True) -> int:
    """Write ``samples`` (an iterable of ``(state, improvements)``) as one shard file; returns the number of records.
    File: int32 magic, int32 0, int64 n, int64 offsets[n + 1] (from the start of the record area, which begins at the
    next multiple of 64 bytes), then the records back to back."""
    records = [pack_sample(s, imp, compress_rows) for s, imp in samples]
    offsets = np.zeros(len(records) + 1, dtype=np.int64)
    np.cumsum([len(r) for r in records], out=offsets[1:])
    head = struct.pack("<iiq", SHARD_MAGIC, 0, len(records)) + offsets.tobytes()
    head += b"\0" * (-len(head) % 64)
    with open(path, "wb") as fh:
        fh.write(head)
        for r in records:
            fh.write(r)
    return len(records)


def convert_sample_files(sample_files, shard_path: str) -> int:
    """Pack the reference's gzip-pickled sample files (data_collector.py:135-140) into one shard."""
    def read(fn):
        with gzip.open(fn, "rb") as fh:
            state, improvements = pickle.load(fh)["data"]
        return state, improvements
    return write_shard(shard_path, (read(fn) for fn in sample_files))


class ShardReader:
    """A shard held in (pinned, when CUDA is available) host memory; ``pointers(ids)`` yields what
    ``GCNN.stage_records`` takes.  Records of consecutive ids are neighbours in memory and travel in one copy."""

    def __init__(self, path: str, pin: bool = True):
        import torch
        raw = np.fromfile(path, dtype=np.uint8)
        magic, _, n = struct.unpack_from("<iiq", raw, 0)
        if magic != SHARD_MAGIC:
            raise ValueError(f"{path} is not a shard file")
        self.offsets = np.frombuffer(raw, dtype=np.int64, count=n + 1, offset=16).copy()
        base = (16 + 8 * (n + 1) + 63) & ~63
        self.buffer = torch.from_numpy(raw[base:].copy())
        if pin and torch.cuda.is_available():
            self.buffer = self.buffer.pin_memory()
        self._view = self.buffer.numpy()
        self._base = self.buffer.data_ptr()
        self._resident = None  # the shard's copy in device memory (to_device)
        self.counts = np.array([record_counts(self._view, int(o))[1:6] for o in self.offsets[:-1]], dtype=np.int64).reshape(n, 5)

    def __len__(self):
        return self.offsets.shape[0] - 1

    @property
    def host_base(self) -> int:
        return self._base

    def to_device(self, device) -> "ShardReader":
        """Keep a copy of the whole shard in the memory of ``device``: ``GCNN.stage_records`` then assembles batches
        straight from it and only ~130 bytes of descriptor per record cross PCIe per step.  (model_trainer.py:147-153
        re-reads every sample file every epoch; a rank's 1/8 of the 100,000-sample training set is ~4 GB of HBM.)"""
        import torch
        self._resident = self.buffer.to(torch.device(device), non_blocking=False)
        torch.cuda.synchronize(self._resident.device)
        return self

    def device_buffer(self, device):
        import torch
        r = self._resident
        return r if r is not None and r.device == torch.device(device) else None

    def pointers(self, ids) -> np.ndarray:
        return (self._base + self.offsets[np.asarray(ids, dtype=np.int64)]).astype(np.uint64)

    def totals(self, ids):
        """(n_cons, n_vars, n_cuts, n_cons_edges, n_cut_edges) of the batch made of records ``ids``."""
        return tuple(int(x) for x in self.counts[np.asarray(ids, dtype=np.int64)].sum(axis=0))

    def sample(self, i: int):
        return unpack_record(self._view, int(self.offsets[i]))

    def record_bytes(self, ids) -> int:
        ids = np.asarray(ids, dtype=np.int64)
        return int((self.offsets[ids + 1] - self.offsets[ids]).sum())


def record_bytes(n_cons, n_vars, n_cuts, n_cons_edges, n_cut_edges, flags) -> int:
    """Record size according to the library (cross-check of the two layout implementations)."""
    return int(_lib.load().gcnn_record_bytes(n_cons, n_vars, n_cuts, n_cons_edges, n_cut_edges, flags))

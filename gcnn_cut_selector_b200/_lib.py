"""ctypes binding of libgcnn_b200.so (include/gcnn_b200.h).  No CPU fallback: a missing library is a hard error."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# GCNN_LIB: an instrumented build (scripts/); GCNN_LIB_VARIANT=alt: build/alt/libgcnn_b200.so, the -DGCNN_ALT_PATHS build
# with the round-1 A/B alternates (python -m gcnn_cut_selector_b200.build -DGCNN_ALT_PATHS --variant=alt)
_VARIANT = os.environ.get("GCNN_LIB_VARIANT")
LIB_PATH = os.environ.get("GCNN_LIB") or (os.path.join(HERE, "build", _VARIANT, "libgcnn_b200.so") if _VARIANT else
                                         os.path.join(HERE, "libgcnn_b200.so"))

OK, INVALID, CUDA_ERROR, OOM = 0, 1, 2, 3
N_TRAINABLE, N_PRENORM, N_ARRAYS, N_PRENORM_LAYERS = 93121, 58, 62, 11


class GcnnError(RuntimeError):
    """CUDA-side failure reported by the library (status GCNN_CUDA_ERROR)."""


class InvalidArgumentError(ValueError):
    """Bad argument or edge index out of range -- what TF-CPU's tf.gather raises in the reference (model.py:564)."""


class ResourceExhaustedError(MemoryError):
    """The batch does not fit on the device.  Stands in for ``tf.errors.ResourceExhaustedError``, which the
    reference's loops catch to skip a batch (model_trainer.py:224-227, 308-311; model_tester.py:229-232)."""


class Batch(C.Structure):
    """``gcnn_batch``: the model input 10-tuple as raw pointers + sizes."""
    _fields_ = [("cons_feats", C.c_void_p), ("cons_edge_inds", C.c_void_p), ("cons_edge_feats", C.c_void_p),
                ("var_feats", C.c_void_p), ("cut_feats", C.c_void_p), ("cut_edge_inds", C.c_void_p),
                ("cut_edge_feats", C.c_void_p), ("n_cons", C.c_int64), ("n_vars", C.c_int64), ("n_cuts", C.c_int64),
                ("n_cons_edges", C.c_int64), ("n_cut_edges", C.c_int64), ("flags", C.c_int64),
                ("sample_n_cons", C.c_void_p), ("sample_n_vars", C.c_void_p), ("sample_n_cuts", C.c_void_p),
                ("n_samples", C.c_int64), ("cons_row_ptr", C.c_void_p), ("cut_row_ptr", C.c_void_p),
                ("cons_col16", C.c_void_p), ("cut_col16", C.c_void_p), ("packed", C.c_void_p), ("packed_bytes", C.c_int64)]


BATCH_CONS_EDGES_SORTED, BATCH_CUT_EDGES_SORTED = 1, 2
MAX_RECORDS = 4096  # samples per batch the staging slots keep block offsets for (csrc/common.cuh)

_P, _I64, _I, _F = C.c_void_p, C.c_int64, C.c_int, C.c_float
_BP = C.POINTER(Batch)

# name -> (restype, argtypes); every symbol include/gcnn_b200.h declares
SIGNATURES = {
    "gcnn_version": (_I, []),
    "gcnn_batch_bytes": (_I64, []),
    "gcnn_last_error": (C.c_char_p, []),
    "gcnn_kernel_launches": (_I, []),
    "gcnn_profile_begin": (_I, []),
    "gcnn_profile_end": (_I, [C.POINTER(C.c_double), C.POINTER(_I64), C.POINTER(C.c_double), _I]),
    "gcnn_profile_num_classes": (_I, []),
    "gcnn_profile_class_name": (C.c_char_p, [_I]),
    "gcnn_param_info": (_I, [_I, C.c_char_p, _I, C.POINTER(_I64), C.POINTER(_I64), C.POINTER(_I), C.POINTER(_I64)]),
    "gcnn_workspace_create": (_I, [C.POINTER(_P)]),
    "gcnn_workspace_destroy": (_I, [_P]),
    "gcnn_workspace_reserve": (_I, [_P, _I64, _I64, _I64, _I64, _I64, _I]),
    "gcnn_workspace_bytes": (_I64, [_P]),
    "gcnn_set_option": (_I, [_P, C.c_char_p, _I]),
    "gcnn_check": (_I, [_P, _P]),
    "gcnn_build_csr": (_I, [_P, _I, _P, _P, _I64, _I64, _I64, _I, _P]),
    "gcnn_build_csr_blocks": (_I, [_P, _I, _P, _P, _I64, _I64, _I64, _P, _P, _I64, _P]),
    "gcnn_csr_export": (_I, [_P, _I, _I, _P, _P, _P, _P, _P]),
    "gcnn_forward": (_I, [_P, _P, _P, _BP, _P, _I, _P]),
    "gcnn_backward": (_I, [_P, _P, _P, _BP, _P, _P, _P]),
    "gcnn_activation_stamp": (_I64, [_P]),
    "gcnn_mse_seed": (_I, [_P, _P, _I64, _F, _P, _P, _P]),
    "gcnn_adam_step": (_I, [_P, _P, _P, _P, _I64, _F, _F, _F, _F, _I64, _P, _P]),
    "gcnn_ranking_deviation": (_I, [_P, _P, _P, _I64, _I, _P, _P]),
    "gcnn_dp_create": (_I, [_P, _I, _I, _P]),
    "gcnn_dp_connect": (_I, [_P, _P]),
    "gcnn_dp_bucket": (_P, [_P, _I]),
    "gcnn_dp_next_parity": (_I, [_P]),
    "gcnn_dp_allreduce_adam": (_I, [_P, _P, _P, _P, _F, _F, _F, _F, _I64, _P, _P]),
    "gcnn_select_cuts": (_I, [_P, _P, _P, _I64, _I64, C.c_double, C.c_double, _I64, _P, _P, _P]),
    "gcnn_forward_backward": (_I, [_P, _P, _P, _BP, _P, _F, _P, _P, _P, _P]),
    "gcnn_prenorm_stats": (_I, [_P, _P, _P, _BP, _I, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                C.POINTER(C.c_double), _P]),
    "gcnn_score_host": (_I, [_P, _P, _P, _BP, _P, _P]),
    "gcnn_score_host_graph": (_I, [_P, _P, _P, _BP, _P, _P]),
    "gcnn_serve_graph_count": (_I, [_P]),
    "gcnn_train_step_host": (_I, [_P, _P, _P, _P, _P, _BP, _P, _F, _I64, C.POINTER(_F), _P]),
    "gcnn_stage_host_batch": (_I, [_P, _I, _BP, _P]),
    "gcnn_record_bytes": (_I64, [_I64, _I64, _I64, _I64, _I64, _I]),
    "gcnn_stage_records": (_I, [_P, _I, _P, _I64, C.POINTER(_I64)]),
    "gcnn_stage_resident_records": (_I, [_P, _I, _P, _P, _P, _I64, C.POINTER(_I64)]),
    "gcnn_score_staged": (_I, [_P, _I, _P, _P, _P, _P]),
    "gcnn_train_step_staged": (_I, [_P, _I, _P, _P, _P, _P, _F, _I64, C.POINTER(_F), _P]),
    "gcnn_train_step_staged_async": (_I, [_P, _I, _P, _P, _P, _P, _F, _I64, _P]),
    "gcnn_dp_train_step_staged_async": (_I, [_P, _I, _P, _P, _P, _P, _F, _I64, _P]),
    "gcnn_train_step_result": (_I, [_P, _I, C.POINTER(_F), _P]),
    "gcnn_staged_batch": (_I, [_P, _I, _BP, C.POINTER(_P), _P]),
    "gcnn_release_staged": (_I, [_P, _I, _P]),
    "gcnn_edge_forward": (_I, [_P, _P, _P, _I64, _P, _P, _P, _F, _F, _F, _P, _P, _P]),
    "gcnn_edge_backward": (_I, [_P, _P, _P, _P, _I64, _P, _P, _P, _P, _F, _F, _F, _P, _P, _P]),
    "gcnn_linear_forward": (_I, [_P, _P, _P, _I64, _I, _I, _P, _P]),
    "gcnn_conv_forward": (_I, [_P, _P, _P, _I, _P, _P, _P, _I64, _P, _P, _P, _P, _P, _P]),
    "gcnn_conv_backward": (_I, [_P, _P, _P, _I, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _P, _P, _P, _P, _P]),
    "gcnn_embed_forward": (_I, [_P, _P, _P, _I, _P, _I64, _P, _P, _P, _P, _P]),
    "gcnn_embed_backward": (_I, [_P, _P, _P, _I, _P, _P, _P, _P, _P, _P, _I64, _P, _P]),
    "gcnn_head_forward": (_I, [_P, _P, _P, _I64, _P, _P]),
    "gcnn_head_backward": (_I, [_P, _P, _P, _P, _I64, _P, _P, _P]),
    "gcnn_has_alt_paths": (_I, []),
}

_lib = None


def load():
    """Load the in-tree shared library and declare every prototype; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m gcnn_cut_selector_b200.build` "
                "(there is no CPU fallback for the GCNN hot path)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the header and the library ever disagree
            fn.restype, fn.argtypes = res, args
        if lib.gcnn_batch_bytes() != C.sizeof(Batch):
            raise ImportError(f"{LIB_PATH}: gcnn_batch is {lib.gcnn_batch_bytes()} bytes in the library but "
                              f"{C.sizeof(Batch)} in this binding (rebuild: python -m gcnn_cut_selector_b200.build)")
        _lib = lib
    return _lib


def check(status: int):
    if status == OK:
        return
    msg = load().gcnn_last_error().decode(errors="replace")
    if status == INVALID:
        raise InvalidArgumentError(msg)
    if status == OOM:
        raise ResourceExhaustedError(msg)
    raise GcnnError(msg)


def param_table():
    """[(name, shape, trainable, offset)] for the 62 arrays in save_state order (model.py:47-56)."""
    lib = load()
    out = []
    buf = C.create_string_buffer(64)
    for i in range(N_ARRAYS):
        rows, cols, off, tr = _I64(), _I64(), _I64(), _I()
        check(lib.gcnn_param_info(i, buf, 64, C.byref(rows), C.byref(cols), C.byref(tr), C.byref(off)))
        shape = (rows.value,) if cols.value == 0 else (rows.value, cols.value)
        out.append((buf.value.decode(), shape, bool(tr.value), off.value))
    return out

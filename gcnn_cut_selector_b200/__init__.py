"""B200-native GCNN message-passing hot path of stefanvanberkum/gcnn-cut-selector (reference: model.py).

Python host code over hand-written sm_100a CUDA kernels (``libgcnn_b200.so``, C ABI in ``include/gcnn_b200.h``).
There is no CPU fallback: constructing ``GCNN`` without the built library or without a CUDA device raises.
"""
from . import batching, shards, synth  # noqa: F401
from ._lib import GcnnError, InvalidArgumentError, ResourceExhaustedError  # noqa: F401
from .batching import load_batch  # noqa: F401
from .metrics import ranking_accuracy, ranking_deviation, select_cuts  # noqa: F401
from .model import GCNN, HostBatch, PreNormException, PreNormLayer, StagedRecords  # noqa: F401
from .trainer import DataParallelTrainer, reduce_bucket  # noqa: F401

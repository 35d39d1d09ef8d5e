"""``GCNN``: host-side mirror of the reference model class (model.py:136-300) over libgcnn_b200.so.

Same constructor, call signature and method surface the reference's drivers use:
``GCNN()``; ``model(batched_states, training)`` with the 10-tuple of model.py:283-284 -> flat fp32 ``[n_cuts]``;
``model.call`` re-assignable, ``model.input_signature``; ``trainable_variables`` / ``variables``;
``save_state`` / ``restore_state`` (the reference's pickle stream, model.py:47-67);
``pretrain_init`` / ``pretrain`` / ``pretrain_next`` (model.py:69-133) with ``PreNormLayer`` semantics
(model.py:384-437).  PyTorch is used for device memory and streams only; all arithmetic runs in the CUDA library.
"""
from __future__ import annotations

import ctypes as C
import os
import pickle
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from ._lib import Batch, InvalidArgumentError, ResourceExhaustedError, check  # noqa: F401

EMB_SIZE, CONS_FEATS, EDGE_FEATS, VAR_FEATS, CUT_FEATS = 64, 4, 1, 14, 6


def _row0_sorted(ei) -> bool:
    """True when a HOST edge-index array has a non-decreasing row 0 (every reference-produced batch does,
    utils.py:102-104, 403-407).  Device tensors are not inspected (that would synchronise)."""
    if torch.is_tensor(ei):
        if ei.is_cuda:
            return False
        ei = ei.numpy()
    ei = np.asarray(ei)
    if ei.ndim != 2 or ei.shape[0] != 2:
        return False
    return bool(ei.shape[1] < 2 or np.all(ei[0, 1:] >= ei[0, :-1]))


def _host(x):
    return x.detach().cpu().numpy() if torch.is_tensor(x) else np.asarray(x)


def _sample_counts(n_cons, n_vars, n_cuts):
    """Three contiguous int32 host vectors when all of n_cons / n_vars / n_cuts are per-sample vectors, else None."""
    vecs = [_host(x) for x in (n_cons, n_vars, n_cuts)]
    if any(v.ndim != 1 for v in vecs) or len({v.shape[0] for v in vecs}) != 1 or vecs[0].shape[0] == 0:
        return None
    return [np.ascontiguousarray(v, dtype=np.int32) for v in vecs]


def _sorted_flags(cons_ei, cut_ei) -> int:
    return ((_lib.BATCH_CONS_EDGES_SORTED if _row0_sorted(cons_ei) else 0)
            | (_lib.BATCH_CUT_EDGES_SORTED if _row0_sorted(cut_ei) else 0))


class PreNormException(Exception):
    """Raised inside ``call`` when an armed pre-norm layer received a batch (model.py:440)."""


@dataclass
class TensorSpec:
    shape: tuple
    dtype: object


class PreNormLayer:
    """Host-side state of one pre-norm layer (model.py:303-437); its shift/scale live in ``GCNN.flat_prenorm``."""

    def __init__(self, model: "GCNN", index: int, name: str, n_units: int, shift_off: int | None, scale_off: int):
        self._model, self.index, self.name, self.n_units = model, index, name, n_units
        self._shift_off, self._scale_off = shift_off, scale_off
        self.waiting_updates = False
        self.received_updates = False
        self.mean = self.var = self.m2 = self.count = None

    @property
    def shift(self):
        if self._shift_off is None:
            return None
        return self._model.flat_prenorm[self._shift_off:self._shift_off + self.n_units]

    @property
    def scale(self):
        return self._model.flat_prenorm[self._scale_off:self._scale_off + self.n_units]

    def start_updates(self):  # model.py:384-392
        self.mean, self.var, self.m2, self.count = 0, 0, 0, 0
        self.waiting_updates, self.received_updates = True, False

    def update_params(self, sample_mean, sample_var, sample_count):
        """Chan parallel merge in fp32, line by line as model.py:416-423."""
        f = np.float32
        sample_mean, sample_var = np.asarray(sample_mean, f), np.asarray(sample_var, f)
        sample_count = f(sample_count)
        delta = sample_mean - self.mean
        self.m2 = self.var * self.count + sample_var * sample_count + delta ** 2 * self.count * sample_count / (
            self.count + sample_count)
        self.count = self.count + sample_count
        self.mean = self.mean + delta * sample_count / self.count
        self.var = self.m2 / self.count if self.count > 0 else 1

    def stop_updates(self):  # model.py:425-437
        mean = np.asarray(self.mean, np.float32).reshape(-1)
        var = np.asarray(self.var, np.float32).reshape(-1)
        if self._shift_off is not None:
            self.shift.copy_(torch.from_numpy(-mean))
        var = np.where(var == 0, np.ones_like(var), var)
        self.scale.copy_(torch.from_numpy((1 / np.sqrt(var)).astype(np.float32)))
        self.mean = self.var = self.m2 = self.count = None
        self.waiting_updates = False


class _GCNNFunction(torch.autograd.Function):
    """Bridges ``tape.gradient`` (model_trainer.py:272) to gcnn_forward / gcnn_backward."""

    @staticmethod
    def forward(ctx, flat_params, model, dev_inputs):
        ctx.model, ctx.dev_inputs = model, dev_inputs
        out = model._forward(dev_inputs, save_activations=True)
        ctx.stamp = model._activation_stamp()  # the workspace's activations are shared: a later forward overwrites them
        return out

    @staticmethod
    def backward(ctx, d_scores):
        if ctx.model._activation_stamp() != ctx.stamp:
            raise RuntimeError("the activations of this forward pass were overwritten by a later call on the same model; "
                               "run backward() before the next forward (or use loss_and_grads)")
        return ctx.model._backward(ctx.dev_inputs, d_scores.contiguous()), None, None


class GCNN:
    """The graph convolutional neural network model (reference: model.py:136-300), B200 implementation."""

    def __init__(self, device=None, seed: int | None = None):
        self._lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("GCNN needs a CUDA device: the hot path has no CPU fallback")
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.name = "gcnn"
        self.emb_size, self.cons_feats, self.edge_feats = EMB_SIZE, CONS_FEATS, EDGE_FEATS
        self.var_feats, self.cut_feats = VAR_FEATS, CUT_FEATS
        self.check_indices = True  # TF-CPU raises on an out-of-range gather index; costs one stream sync per call

        self._table = _lib.param_table()
        self.flat_params = torch.zeros(_lib.N_TRAINABLE, dtype=torch.float32, device=self.device, requires_grad=True)
        self.flat_prenorm = torch.zeros(_lib.N_PRENORM, dtype=torch.float32, device=self.device)
        self.flat_grads = torch.zeros(_lib.N_TRAINABLE, dtype=torch.float32, device=self.device)
        self._init_weights(seed)

        self.variables_topological_order = [name for name, _, _, _ in self._table]  # model.py:215
        f32, i32 = torch.float32, torch.int32
        self.input_signature = [(TensorSpec((None, CONS_FEATS), f32), TensorSpec((2, None), i32),
                                 TensorSpec((None, EDGE_FEATS), f32), TensorSpec((None, VAR_FEATS), f32),
                                 TensorSpec((None, CUT_FEATS), f32), TensorSpec((2, None), i32),
                                 TensorSpec((None, EDGE_FEATS), f32), TensorSpec((), i32), TensorSpec((), i32),
                                 TensorSpec((), i32)), TensorSpec((), torch.bool)]  # model.py:218-226

        ws = C.c_void_p()
        with torch.cuda.device(self.device):  # the workspace (streams, events, arena) lives on the model's device
            check(self._lib.gcnn_workspace_create(C.byref(ws)))
        self._ws = ws
        self._params_seen, self._params_epoch = None, 0
        self._prenorm_layers = self._make_prenorm_layers()
        self.call = self._call  # re-assignable like ``model.call = tf.function(model.call, ...)`` (model_trainer.py:144)

        self._staged = [None, None]
        # optimiser state for the fused train step (Keras Adam, model_trainer.py:131)
        self.adam_m = torch.zeros_like(self.flat_grads)
        self.adam_v = torch.zeros_like(self.flat_grads)
        self.adam_step = 0
        self._loss2 = torch.zeros(2, dtype=torch.float32, device=self.device)  # [cut count (optional) | loss sum]
        self._loss_sum = self._loss2[1:2]

    def __del__(self):
        try:
            if getattr(self, "_ws", None):
                self._lib.gcnn_workspace_destroy(self._ws)
                self._ws = None
        except Exception:
            pass

    # ---- parameters ------------------------------------------------------------------------------------------------
    def _view(self, entry):
        name, shape, trainable, off = entry
        n = int(np.prod(shape))
        base = self.flat_params.detach() if trainable else self.flat_prenorm
        return base[off:off + n].view(*shape)

    @property
    def variables(self):
        """All 62 arrays as views of the flat buffers, in the reference's ``model.variables`` order."""
        return [self._view(e) for e in self._table]

    @property
    def trainable_variables(self):
        return [self._view(e) for e in self._table if e[2]]

    @property
    def trainable_gradients(self):
        out = []
        for name, shape, trainable, off in self._table:
            if trainable:
                out.append(self.flat_grads[off:off + int(np.prod(shape))].view(*shape))
        return out

    def _init_weights(self, seed):
        """Keras defaults of the reference: orthogonal kernels, zero biases (model.py:175), shift 0 / scale 1
        (model.py:334, 342)."""
        # seed=None draws from torch's default generator (like the reference draws from the framework's global seed):
        # reproducible under torch.manual_seed and without reseeding the caller's generator
        gen = torch.Generator().manual_seed(int(seed) if seed is not None
                                            else int(torch.randint(0, 2 ** 31 - 1, (1,)).item()))
        flat = torch.zeros(_lib.N_TRAINABLE)
        for name, shape, trainable, off in self._table:
            if trainable and name.endswith("kernel"):
                big = max(shape)
                q, r = torch.linalg.qr(torch.randn(big, big, generator=gen))
                q = q * torch.sign(torch.diagonal(r))
                flat[off:off + shape[0] * shape[1]] = q[:shape[0], :shape[1]].reshape(-1)
        pn = torch.zeros(_lib.N_PRENORM)
        for name, shape, trainable, off in self._table:
            if not trainable and name.endswith("scale"):
                pn[off:off + shape[0]] = 1.0
        with torch.no_grad():
            self.flat_params.copy_(flat)
            self.flat_prenorm.copy_(pn)

    def save_state(self, path: str):
        """model.py:47-56: one ``pickle.dump(ndarray)`` per variable in ``variables_topological_order``."""
        with open(path, "wb") as fh:
            for v in self.variables:
                pickle.dump(v.detach().cpu().numpy(), fh)

    def restore_state(self, path: str):
        """model.py:58-67.  The stream is positional; a shape mismatch fails loudly."""
        with open(path, "rb") as fh, torch.no_grad():
            for (name, shape, _, _), v in zip(self._table, self.variables):
                arr = np.asarray(pickle.load(fh), dtype=np.float32)
                if tuple(arr.shape) != tuple(shape):
                    raise ValueError(f"weights stream mismatch at {name}: got {arr.shape}, expected {shape}")
                v.copy_(torch.from_numpy(arr))

    # ---- inputs ---------------------------------------------------------------------------------------------------
    def _to_device(self, x, dtype):
        if not torch.is_tensor(x):
            x = torch.from_numpy(np.ascontiguousarray(x))
        return x.to(device=self.device, dtype=dtype, non_blocking=True).contiguous()

    def prepare_inputs(self, inputs):
        """Move the 10-tuple to the device (fp32 / int32, contiguous) and wrap it as a ``gcnn_batch``."""
        (cons, cons_ei, cons_ef, var, cut, cut_ei, cut_ef, n_cons, n_vars, n_cuts) = inputs
        f32, i32 = torch.float32, torch.int32
        flags = _sorted_flags(cons_ei, cut_ei)  # host arrays only: a promise the device still verifies
        t = [self._to_device(cons, f32), self._to_device(cons_ei, i32), self._to_device(cons_ef, f32),
             self._to_device(var, f32), self._to_device(cut, f32), self._to_device(cut_ei, i32),
             self._to_device(cut_ef, f32)]
        # n_cons / n_vars / n_cuts: the totals the reference passes (model_trainer.py:259-263), or the per-sample
        # vectors load_batch returns (utils.py:420-422) -- the latter give the library the batch's block structure:
        # per-sample gathered tables in shared memory and per-sample transposed layouts (csrc/edge_block.cu)
        counts = _sample_counts(n_cons, n_vars, n_cuts)
        n_cons, n_vars, n_cuts = (int(np.sum(_host(x))) for x in (n_cons, n_vars, n_cuts))
        if t[0].shape != (n_cons, CONS_FEATS) and not (n_cons == 0 and t[0].numel() == 0):
            raise InvalidArgumentError(f"cons_feats {tuple(t[0].shape)} vs n_cons {n_cons}")
        if t[3].numel() != n_vars * VAR_FEATS or t[4].numel() != n_cuts * CUT_FEATS:
            raise InvalidArgumentError("var_feats / cut_feats do not match n_vars / n_cuts")
        if t[1].dim() != 2 or t[1].shape[0] != 2 or t[5].dim() != 2 or t[5].shape[0] != 2:
            raise InvalidArgumentError("edge indices must be [2, E]")
        e_c, e_k = t[1].shape[1], t[5].shape[1]
        if t[2].numel() != e_c or t[6].numel() != e_k:
            raise InvalidArgumentError("edge features must be [E, 1]")
        b = Batch(t[0].data_ptr(), t[1].data_ptr(), t[2].data_ptr(), t[3].data_ptr(), t[4].data_ptr(),
                  t[5].data_ptr(), t[6].data_ptr(), n_cons, n_vars, n_cuts, e_c, e_k, flags)
        if counts is not None:
            b.sample_n_cons, b.sample_n_vars, b.sample_n_cuts = (c.ctypes.data for c in counts)
            b.n_samples = counts[0].shape[0]
            t = t + list(counts)  # keep the host arrays alive as long as the batch
        return b, t

    def reserve(self, batch: Batch, training: bool):
        """Grow the workspace for a batch of these sizes.  Batches already staged with ``stage_host`` / ``stage_records``
        survive the growth (the staging slots have their own allocation)."""
        check(self._lib.gcnn_workspace_reserve(self._ws, batch.n_cons, batch.n_vars, batch.n_cuts, batch.n_cons_edges,
                                               batch.n_cut_edges, int(training)))

    def set_option(self, name: str, value: int):
        """Library options (include/gcnn_b200.h): "tensor_cores", "streams", "blocks", ..."""
        check(self._lib.gcnn_set_option(self._ws, name.encode(), int(value)))

    def _stream(self):
        self._announce_params()  # every library call that takes a stream also reads the parameters
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _announce_params(self):
        """Option "params_epoch" (include/gcnn_b200.h): a new epoch whenever the parameter tensor was written through
        torch (its version counter counts in-place writes, also through ``trainable_variables`` views: ``restore_state``,
        ``v.assign``-style copies, an external optimiser) or replaced, so the library re-packs its weight images only
        then -- scoring with frozen weights (model_benchmarker.py:91-106) skips the re-pack.  The library's own update
        calls invalidate the images themselves; ``apply_gradients`` (gcnn_adam_step, no workspace) bumps the epoch here.
        Writes that bypass torch (a raw pointer handed to another library) need ``params_changed()``."""
        seen = (self.flat_params.data_ptr(), self.flat_params._version)
        if seen != self._params_seen:
            self._params_seen = seen
            self._params_epoch = self._params_epoch % (2 ** 31 - 2) + 1
            check(self._lib.gcnn_set_option(self._ws, b"params_epoch", self._params_epoch))

    def params_changed(self):
        """Tell the library that ``flat_params`` was written behind torch's back (see ``_announce_params``)."""
        self._params_seen = None

    def _activation_stamp(self) -> int:
        return int(self._lib.gcnn_activation_stamp(self._ws))

    def _check_indices(self):
        """TF-CPU raises InvalidArgument on an out-of-range gather index (model.py:564); here the kernels clamp and set
        a sticky device word, read back (one stream synchronisation) when ``check_indices`` is on."""
        if self.check_indices:
            check(self._lib.gcnn_check(self._ws, self._stream()))

    # ---- forward / backward ----------------------------------------------------------------------------------------
    def _forward(self, dev_inputs, save_activations: bool):
        batch, _keep = dev_inputs
        self.reserve(batch, save_activations)
        scores = torch.empty(batch.n_cuts, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            check(self._lib.gcnn_forward(self._ws, self.flat_params.data_ptr(), self.flat_prenorm.data_ptr(),
                                         C.byref(batch), scores.data_ptr(), int(save_activations), self._stream()))
            self._check_indices()
        return scores

    def _backward(self, dev_inputs, d_scores):
        batch, _keep = dev_inputs
        with torch.cuda.device(self.device):
            check(self._lib.gcnn_backward(self._ws, self.flat_params.data_ptr(), self.flat_prenorm.data_ptr(),
                                          C.byref(batch), d_scores.data_ptr(), self.flat_grads.data_ptr(), self._stream()))
        return self.flat_grads.clone()

    def _first_armed(self):
        for layer in self._prenorm_layers:
            if layer.waiting_updates:
                return layer
        return None

    def _call(self, inputs, training=False):
        """GCNN.call (model.py:257-300).  ``training`` only selects whether activations are kept for a backward
        pass (the reference threads it through but has no dropout / batch-norm)."""
        dev_inputs = self.prepare_inputs(inputs)
        armed = self._first_armed()
        if armed is not None:  # model.py:372-375: the first armed layer updates its statistics and aborts the call
            self._update_prenorm(armed, dev_inputs)
            armed.received_updates = True
            raise PreNormException
        training = bool(training)
        if training and torch.is_grad_enabled() and self.flat_params.requires_grad:
            return _GCNNFunction.apply(self.flat_params, self, dev_inputs)
        return self._forward(dev_inputs, save_activations=False)

    def __call__(self, inputs, training=False):
        return self.call(inputs, training)

    def get_concrete_function(self):
        """``model.call.get_concrete_function()`` shim (model_benchmarker.py:309-310): returns the bound callable."""
        return self.call

    # ---- training step (model_trainer.py:269-273) -----------------------------------------------------------------
    def loss_and_grads(self, inputs, targets, seed_scale: float | None = None, loss_out: torch.Tensor | None = None):
        """forward + MeanSquaredError + tape.gradient in one library call.  Returns (loss_sum tensor [1], scores);
        gradients land in ``flat_grads``.  ``seed_scale`` defaults to 1/n_cuts (mean over all cuts of the batch).
        ``loss_out``: a one-element fp32 device tensor to receive the loss sum instead of the internal buffer (with
        option "count_before_loss" the element before it receives the cut count)."""
        dev_inputs = inputs if isinstance(inputs, tuple) and isinstance(inputs[0], Batch) else self.prepare_inputs(inputs)
        batch, _keep = dev_inputs
        targets = self._to_device(targets, torch.float32)
        self.reserve(batch, True)
        scores = torch.empty(batch.n_cuts, dtype=torch.float32, device=self.device)
        scale = (1.0 / max(batch.n_cuts, 1)) if seed_scale is None else float(seed_scale)
        with torch.cuda.device(self.device):
            check(self._lib.gcnn_forward_backward(self._ws, self.flat_params.data_ptr(), self.flat_prenorm.data_ptr(),
                                                  C.byref(batch), targets.data_ptr(), scale, scores.data_ptr(),
                                                  self.flat_grads.data_ptr(),
                                                  (loss_out if loss_out is not None else self._loss_sum).data_ptr(),
                                                  self._stream()))
            self._check_indices()
        return (loss_out if loss_out is not None else self._loss_sum), scores

    def apply_gradients(self, lr: float, grad_divisor: torch.Tensor | None = None,
                        beta1=0.9, beta2=0.999, eps=1e-7):
        """Keras ``Adam.apply_gradients`` on the flat buffers (one launch instead of 46)."""
        div = grad_divisor.data_ptr() if grad_divisor is not None else None
        with torch.cuda.device(self.device):
            check(self._lib.gcnn_adam_step(self.flat_params.data_ptr(), self.flat_grads.data_ptr(),
                                           self.adam_m.data_ptr(), self.adam_v.data_ptr(), _lib.N_TRAINABLE, lr, beta1,
                                           beta2, eps, self.adam_step + 1, div, self._stream()))
        self.adam_step += 1  # only a step that was enqueued counts (Adam's bias correction depends on it)
        self._params_seen = None  # gcnn_adam_step wrote the parameters without the workspace knowing

    def train_step(self, inputs, targets, lr: float):
        """One optimisation step on device-resident or host inputs; returns the mean loss as a device tensor."""
        loss_sum, scores = self.loss_and_grads(inputs, targets)
        self.apply_gradients(lr)
        return loss_sum / max(scores.numel(), 1), scores

    def train_step_host(self, host_batch: "HostBatch", lr: float) -> float:
        """End-to-end step from (pinned) host buffers: H2D copies, forward, MSE, backward, Adam, loss back to host."""
        b = host_batch.batch
        self.reserve(b, True)
        loss = C.c_float()
        with torch.cuda.device(self.device):
            check(self._lib.gcnn_train_step_host(self._ws, self.flat_params.data_ptr(), self.flat_prenorm.data_ptr(),
                                                 self.adam_m.data_ptr(), self.adam_v.data_ptr(), C.byref(b),
                                                 host_batch.targets.data_ptr(), lr, self.adam_step + 1, C.byref(loss),
                                                 self._stream()))
        self.adam_step += 1
        return float(loss.value)

    # ---- prefetching host path: batch i + 1 is copied in while the step on batch i runs (model_trainer.py:153) ------
    def stage_host(self, host_batch: "HostBatch", slot: int, training: bool = True):
        """Enqueue the host-to-device copies of ``host_batch`` into staging slot 0/1 (returns immediately)."""
        b = host_batch.batch
        self.reserve(b, training)
        tgt = host_batch.targets.data_ptr() if training else None
        with torch.cuda.device(self.device):
            check(self._lib.gcnn_stage_host_batch(self._ws, slot, C.byref(b), tgt))
        self._staged[slot] = host_batch  # keeps the pinned buffers alive until the slot is consumed

    def stage_records(self, reader, ids, slot: int, training: bool = True) -> "StagedRecords":
        """Stage the batch made of records ``ids`` of a ``shards.ShardReader`` in slot 0/1: the packed records are copied
        to the device as they are and one kernel assembles the batch there (the device-side ``utils.load_batch``,
        utils.py:339-426).  If the reader's shard is resident on this model's device (``ShardReader.to_device``) no
        record bytes are copied at all.  Returns immediately; use the slot like one filled by ``stage_host``."""
        nc, nv, nk, ec, ek = reader.totals(ids)
        ptrs = reader.pointers(ids)
        h2d = C.c_int64()
        resident = reader.device_buffer(self.device)  # set by ShardReader.to_device: the shard lives in HBM
        with torch.cuda.device(self.device):
            check(self._lib.gcnn_workspace_reserve(self._ws, nc, nv, nk, ec, ek, int(training)))
            if resident is not None:
                check(self._lib.gcnn_stage_resident_records(self._ws, slot, resident.data_ptr(), reader.host_base,
                                                            ptrs.ctypes.data, ptrs.shape[0], C.byref(h2d)))
            else:
                check(self._lib.gcnn_stage_records(self._ws, slot, ptrs.ctypes.data, ptrs.shape[0], C.byref(h2d)))
        staged = StagedRecords(reader, nk, int(ptrs.shape[0]), int(h2d.value))
        self._staged[slot] = staged  # keeps the pinned shard alive until the slot is consumed
        return staged

    def train_step_staged(self, slot: int, lr: float) -> float:
        """Optimisation step on the batch staged in ``slot``; returns the mean loss (host float)."""
        loss = C.c_float()
        with torch.cuda.device(self.device):
            check(self._lib.gcnn_train_step_staged(self._ws, slot, self.flat_params.data_ptr(),
                                                   self.flat_prenorm.data_ptr(), self.adam_m.data_ptr(),
                                                   self.adam_v.data_ptr(), lr, self.adam_step + 1, C.byref(loss),
                                                   self._stream()))
        self.adam_step += 1
        return float(loss.value)

    def train_step_staged_async(self, slot: int, lr: float):
        """Enqueue the optimisation step on the batch staged in ``slot`` without waiting for it; its mean loss is read
        later with ``train_step_result(slot)`` (e.g. after the next step has been enqueued)."""
        with torch.cuda.device(self.device):
            check(self._lib.gcnn_train_step_staged_async(self._ws, slot, self.flat_params.data_ptr(),
                                                         self.flat_prenorm.data_ptr(), self.adam_m.data_ptr(),
                                                         self.adam_v.data_ptr(), lr, self.adam_step + 1, self._stream()))
        self.adam_step += 1

    def train_step_result(self, slot: int) -> float:
        loss = C.c_float()
        with torch.cuda.device(self.device):
            check(self._lib.gcnn_train_step_result(self._ws, slot, C.byref(loss), self._stream()))
        return float(loss.value)

    def loss_and_grads_staged(self, slot: int, seed_scale: float | None = None, loss_out: torch.Tensor | None = None):
        """``loss_and_grads`` on the batch staged in ``slot`` (data-parallel trainer).  Returns (loss_sum, n_cuts)."""
        batch, tgt = Batch(), C.c_void_p()
        with torch.cuda.device(self.device):
            check(self._lib.gcnn_staged_batch(self._ws, slot, C.byref(batch), C.byref(tgt), self._stream()))
            scale = (1.0 / max(batch.n_cuts, 1)) if seed_scale is None else float(seed_scale)
            check(self._lib.gcnn_forward_backward(self._ws, self.flat_params.data_ptr(), self.flat_prenorm.data_ptr(),
                                                  C.byref(batch), tgt, scale, None, self.flat_grads.data_ptr(),
                                                  (loss_out if loss_out is not None else self._loss_sum).data_ptr(),
                                                  self._stream()))
            check(self._lib.gcnn_release_staged(self._ws, slot, self._stream()))
            self._check_indices()
        return (loss_out if loss_out is not None else self._loss_sum), int(batch.n_cuts)

    def score_staged(self, slot: int) -> np.ndarray:
        """Cut scores of the batch staged in ``slot`` (inference), as a host array."""
        out = self._staged[slot].scores
        with torch.cuda.device(self.device):
            check(self._lib.gcnn_score_staged(self._ws, slot, self.flat_params.data_ptr(), self.flat_prenorm.data_ptr(),
                                              out.data_ptr(), self._stream()))
        return out.numpy()

    def score_host(self, host_batch: "HostBatch", graph: bool = False) -> np.ndarray:
        """Cut scoring from host buffers to a host array (the ``get_improvements(state, False).numpy()`` path of
        model_benchmarker.py:106).  ``graph=True``: the whole call replays as one CUDA graph once its shape has been seen
        twice (the plug-in's one-graph-per-call loop; include/gcnn_b200.h ``gcnn_score_host_graph``)."""
        b = host_batch.batch
        self.reserve(b, False)
        out = host_batch.scores
        fn = self._lib.gcnn_score_host_graph if graph else self._lib.gcnn_score_host
        with torch.cuda.device(self.device):
            check(fn(self._ws, self.flat_params.data_ptr(), self.flat_prenorm.data_ptr(), C.byref(b), out.data_ptr(),
                     self._stream()))
        return out.numpy()

    # ---- pre-norm pretraining (model.py:69-133) -------------------------------------------------------------------
    def _make_prenorm_layers(self):
        offs = {name: off for name, _, _, off in self._table}
        spec = [("cons_emb/prenorm", CONS_FEATS, True), ("cons_edge/prenorm", 1, True),
                ("var_emb/prenorm", VAR_FEATS, True), ("cut_emb/prenorm", CUT_FEATS, True),
                ("cut_edge/prenorm", 1, True)]
        for conv in ("cons_conv", "var_conv", "cut_conv"):
            spec += [(f"{conv}_final/prenorm", 1, False), (f"{conv}_post/prenorm", 1, False)]
        return [PreNormLayer(self, i, name, n, offs[name + "/shift"] if has_shift else None, offs[name + "/scale"])
                for i, (name, n, has_shift) in enumerate(spec)]

    def _prenorm_batch_stats(self, layer: PreNormLayer, dev_inputs):
        """(mean, population variance, count) of the input of ``layer`` on this batch (model.py:410-413), on the host."""
        batch, _keep = dev_inputs
        self.reserve(batch, False)
        mean = (C.c_double * 64)()
        var = (C.c_double * 64)()
        count = C.c_double()
        with torch.cuda.device(self.device):
            check(self._lib.gcnn_prenorm_stats(self._ws, self.flat_params.data_ptr(), self.flat_prenorm.data_ptr(),
                                               C.byref(batch), layer.index, mean, var, C.byref(count), self._stream()))
        n = layer.n_units
        return np.array(mean[:n]), np.array(var[:n]), count.value

    def _update_prenorm(self, layer: PreNormLayer, dev_inputs):
        layer.update_params(*self._prenorm_batch_stats(layer, dev_inputs))

    def pretrain_init(self):
        for layer in self._prenorm_layers:
            layer.start_updates()

    def pretrain_next(self):
        for layer in self._prenorm_layers:
            if layer.waiting_updates and layer.received_updates:
                layer.stop_updates()
                return layer, f"{self.name}/{layer.name}"
        return None

    def pretrain_fused(self, batches) -> int:
        """The whole pre-norm pretraining of ``model_trainer.pretrain`` (model_trainer.py:194-236) in 7 passes over
        ``batches`` (a re-iterable of model input tuples) instead of 11: the five input layers normalise raw features
        (model.py:174-198), so their statistics depend on no parameter and share the first pass; the six convolution
        layers form a chain (each needs everything before it frozen, model.py:100-117) and keep one pass each.  Every
        layer sees the same batches in the same order as in the reference's loop, so the Chan merges (model.py:416-423)
        and the frozen shift / scale values are identical to the 11-pass protocol.  Returns the number of passes."""
        self.pretrain_init()
        layers, passes = self._prenorm_layers, 0
        for group in [layers[:5]] + [[layer] for layer in layers[5:]]:
            seen = False
            for b in batches:
                dev_inputs = self.prepare_inputs(b)
                for layer in group:
                    self._update_prenorm(layer, dev_inputs)
                    layer.received_updates = True
                seen = True
            if not seen:
                break
            for layer in group:
                layer.stop_updates()
            passes += 1
        return passes

    def pretrain(self, *args, **kwargs) -> bool:
        try:
            with torch.no_grad():
                self.call(*args, **kwargs)
            return False
        except PreNormException:
            return True


class StagedRecords:
    """What ``GCNN.stage_records`` put into a staging slot (keeps the shard's pinned memory alive)."""

    def __init__(self, reader, n_cuts: int, n_graphs: int, h2d_bytes: int):
        self.reader, self.n_cuts, self.n_graphs, self.h2d_bytes = reader, n_cuts, n_graphs, h2d_bytes
        self._scores = None

    @property
    def scores(self):  # pinned result buffer of score_staged; allocated on first use (training never needs it)
        if self._scores is None:
            self._scores = torch.empty(self.n_cuts, dtype=torch.float32).pin_memory()
        return self._scores


class HostBatch:
    """A batch in pinned host memory plus its ``gcnn_batch`` of HOST pointers, for the ``*_host`` entry points."""

    ROW_POINTER_MIN_EDGES = 1 << 17  # below this the two extra copies + expansions cost more host time than the bytes save

    def __init__(self, batch11, row_pointers: bool | None = None, packed: bool = True):
        (cons, cons_ei, cons_ef, var, cut, cut_ei, cut_ef, n_cons, n_vars, n_cuts, targets) = batch11
        arr = lambda a, dt: np.ascontiguousarray(np.asarray(a, dtype=dt))
        host = [arr(cons, np.float32), arr(cons_ei, np.int32), arr(cons_ef, np.float32), arr(var, np.float32),
                arr(cut, np.float32), arr(cut_ei, np.int32), arr(cut_ef, np.float32)]
        tgt = arr(targets, np.float32)
        nc, nv, nk = int(np.sum(n_cons)), int(np.sum(n_vars)), int(np.sum(n_cuts))
        flags = _sorted_flags(torch.from_numpy(host[1]), torch.from_numpy(host[5]))
        # sorted edge lists travel as row pointers: 4 of their 12 bytes per edge stay on the host (gcnn_batch::*_row_ptr).
        # row_pointers=None: only for large lists -- measured on one box (profiles/r2_ab_host_paths.json): a 32-graph
        # set-cover step 0.436 -> 0.428 ms end to end, but +25-30 us on a single-graph scoring / training call
        row_ptrs = [None, None]
        for i, (ei, n_rows, flag) in enumerate(((host[1], nc, _lib.BATCH_CONS_EDGES_SORTED), (host[5], nk, _lib.BATCH_CUT_EDGES_SORTED))):
            want = ei.shape[1] >= self.ROW_POINTER_MIN_EDGES if row_pointers is None else bool(row_pointers)
            if row_pointers is None and os.environ.get("GCNN_HOST_ROW_POINTERS") in ("0", "1"):  # A/B switch
                want = os.environ["GCNN_HOST_ROW_POINTERS"] == "1"
            if want and (flags & flag) and n_rows > 0 and ei.shape[1] > 0:
                row_ptrs[i] = np.searchsorted(ei[0], np.arange(n_rows + 1, dtype=np.int64), side="left").astype(np.int32)
        # ... and the column (variable) indices of those lists as uint16 local to the sample that owns the edge's row
        # (gcnn_batch::*_col16): 2 instead of 4 bytes per edge; skipped when a local index does not fit or the edge
        # leaves its sample's block (the library's own checks then see the full indices)
        self.counts = _sample_counts(n_cons, n_vars, n_cuts)
        col16 = [None, None]
        if self.counts is not None and self.counts[0].shape[0] <= _lib.MAX_RECORDS:
            var_off = np.concatenate(([0], np.cumsum(self.counts[1], dtype=np.int64)))
            for i, (ei, left_counts) in enumerate(((host[1], self.counts[0]), (host[5], self.counts[2]))):
                left_end = np.cumsum(left_counts, dtype=np.int64)
                if row_ptrs[i] is None or left_end[-1] != (nc, nk)[i] or var_off[-1] != nv:
                    continue
                sidx = np.searchsorted(left_end, ei[0], side="right").clip(max=len(left_end) - 1)
                local = ei[1] - var_off[sidx]
                if local.size and local.min() >= 0 and local.max() < 65536 and bool(np.all(local < self.counts[1][sidx])):
                    col16[i] = local.astype(np.uint16)
        # one pinned buffer for everything that crosses PCIe whole (gcnn_batch::packed): a single transfer per batch
        # instead of ~13; an index tensor that travels as a row pointer stays outside it (only its columns are copied, or
        # nothing at all with local columns)
        inside = [True, row_ptrs[0] is None, True, True, True, row_ptrs[1] is None, True]
        sections = [(a, k) for k, a in enumerate(host) if inside[k]] + [(tgt, "t")]
        sections += [(a, ("p", i)) for i, a in enumerate(row_ptrs) if a is not None]
        sections += [(a, ("c", i)) for i, a in enumerate(col16) if a is not None]
        pin = lambda a: torch.from_numpy(a).pin_memory()
        placed = {}
        if packed:
            offs, total = [], 0
            for a, _ in sections:
                offs.append(total)
                total += (a.nbytes + 255) & ~255  # (the device copy keeps the staging buffers' 256-byte alignment)
            self.arena = torch.empty(max(total, 16), dtype=torch.uint8).pin_memory()
            for (a, key), off in zip(sections, offs):
                view = self.arena[off:off + a.nbytes].view(torch.from_numpy(a).dtype).view(a.shape)
                view.copy_(torch.from_numpy(a))
                placed[key] = view
        else:
            self.arena = None
            for a, key in sections:
                placed[key] = pin(a)
        self.tensors = [placed[k] if inside[k] else pin(host[k]) for k in range(7)]
        self.targets = placed["t"]
        self.row_ptrs = [placed.get(("p", i)) for i in range(2)]
        self.col16 = [placed.get(("c", i)) for i in range(2)]
        self.scores = torch.empty(nk, dtype=torch.float32).pin_memory()
        t = self.tensors
        self.batch = b = Batch(t[0].data_ptr(), t[1].data_ptr(), t[2].data_ptr(), t[3].data_ptr(), t[4].data_ptr(),
                               t[5].data_ptr(), t[6].data_ptr(), nc, nv, nk, t[1].shape[1], t[5].shape[1], flags)
        if self.row_ptrs[0] is not None:
            b.cons_row_ptr = self.row_ptrs[0].data_ptr()
        if self.row_ptrs[1] is not None:
            b.cut_row_ptr = self.row_ptrs[1].data_ptr()
        if self.col16[0] is not None:
            b.cons_col16 = self.col16[0].data_ptr()
        if self.col16[1] is not None:
            b.cut_col16 = self.col16[1].data_ptr()
        if self.counts is not None:
            b.sample_n_cons, b.sample_n_vars, b.sample_n_cuts = (c.ctypes.data for c in self.counts)
            b.n_samples = self.counts[0].shape[0]
        if self.arena is not None:
            b.packed, b.packed_bytes = self.arena.data_ptr(), total
        self.n_graphs = int(np.size(n_cons))
        # bytes that cross PCIe per staging call: the packed sections (features, coefficients, targets, pointers, local
        # columns, whole index tensors of unsorted lists) + the column row of a list with a pointer but no local columns
        self.h2d_bytes = total if packed else sum(a.nbytes for a, _ in sections)
        for i, k in enumerate((1, 5)):
            if row_ptrs[i] is not None and col16[i] is None:
                self.h2d_bytes += host[k].shape[1] * 4

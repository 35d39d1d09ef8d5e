"""Torch-backed stand-in for the few ``tensorflow`` symbols the reference's model.py / utils.py touch.

TEST INFRASTRUCTURE ONLY (see oracle/gcnn_oracle.py header).  TensorFlow 2.7.1 is not installable in the build
container, so ``oracle/make_golden.py`` puts this directory on ``sys.path`` and imports the reference's own,
unmodified sources over it.  Every op restates TF's documented semantics on torch CPU tensors; the float dtype
behind ``tf.float32`` is switchable so the same reference code can be run in fp64 for ground truth.
"""
import numpy as np
import torch

_FLOAT = torch.float32


def set_float_dtype(dtype):
    global _FLOAT, float32
    _FLOAT = dtype
    float32 = dtype


float32 = torch.float32
int32 = torch.int32
bool = torch.bool  # noqa: A001  (mirrors tf.bool)


class TensorSpec:
    def __init__(self, shape, dtype):
        self.shape, self.dtype = shape, dtype


class _Errors:
    class ResourceExhaustedError(Exception):
        pass


errors = _Errors()


class _Data:  # only named in type annotations of model_trainer.py (tf.data.Dataset)
    class Dataset:
        pass


data = _Data()


def _np(self):
    return self.detach().cpu().numpy()


torch.Tensor.numpy_tf = _np  # not used by the reference; kept for debugging

# A TF eager tensor gives up its value with .numpy() whether or not a tape is watching (model_trainer.py:285-286, 303 do
# that with predictions and the loss inside the training branch); a torch tensor that requires grad refuses.  In this
# process -- the golden generator and the tests that import the shim -- numpy() detaches first.
_torch_numpy = torch.Tensor.numpy


def _eager_numpy(self, *args, **kwargs):
    return _torch_numpy(self.detach().cpu(), *args, **kwargs)


torch.Tensor.numpy = _eager_numpy


def convert_to_tensor(value, dtype=None):
    if dtype is float32 or dtype is torch.float32 or dtype is torch.float64:
        dtype = _FLOAT
    if torch.is_tensor(value):
        return value.to(dtype) if dtype is not None else value
    arr = np.asarray(value)
    if dtype is None and arr.dtype.kind == "f":
        dtype = _FLOAT
    return torch.as_tensor(arr).to(dtype) if dtype is not None else torch.as_tensor(arr)


def gather(params, indices, axis=0):
    idx = indices.long()
    if ((idx < 0) | (idx >= params.shape[axis])).any():  # TF-CPU raises InvalidArgumentError
        raise IndexError("gather index out of range")
    return torch.index_select(params, axis, idx)


def scatter_nd(indices, updates, shape):
    """Zero-initialised tensor of ``shape``; duplicate indices accumulate (sequentially, in update order, on CPU)."""
    shape = [int(s) for s in shape]
    idx = indices.long()
    assert idx.shape[-1] == 1, "shim only supports row scatter"
    return torch.zeros(shape, dtype=updates.dtype).index_add_(0, idx[:, 0], updates)


def expand_dims(x, axis):
    return x.unsqueeze(axis)


def concat(values, axis):
    return torch.cat(list(values), dim=axis)


class Variable:
    """Minimal ``tf.Variable``: a named torch leaf with ``assign`` / ``numpy``."""

    def __init__(self, value, name, trainable):
        self._t = value.detach().clone().requires_grad_(trainable)
        self.name, self.trainable = name, trainable

    @property
    def shape(self):
        return tuple(self._t.shape)

    def assign(self, value):
        value = _val(value)
        if not torch.is_tensor(value):
            value = torch.as_tensor(np.asarray(value))
        with torch.no_grad():
            self._t.copy_(value.to(self._t.dtype).reshape(self._t.shape))

    def numpy(self):
        return self._t.detach().cpu().numpy()


def _val(x):
    return x._t if isinstance(x, Variable) else x


def add(a, b):
    return _val(a) + _val(b)


def multiply(a, b):
    return _val(a) * _val(b)


def reshape(x, shape):
    return x.reshape([int(s) for s in shape])


def reduce_mean(x, axis=None):
    return x.mean() if axis is None else x.mean(axis)


def reduce_sum(x, axis=None):
    return x.sum() if axis is None else x.sum(axis)


def size(input):  # noqa: A002
    return torch.tensor(input.numel())


def cast(x, dtype):
    if dtype is float32 or dtype is torch.float32:
        dtype = _FLOAT
    return torch.as_tensor(x).to(dtype)


def where(cond, a, b):
    return torch.where(cond, a, b)


def equal(a, b):
    return torch.as_tensor(a) == b


def ones_like(x):
    return torch.ones_like(torch.as_tensor(x))


def split(value, num_or_size_splits, axis=0):
    """tf.split with a 1-D tensor of sizes (model_trainer.py:280-281)."""
    sizes = [int(s) for s in (num_or_size_splits.tolist() if torch.is_tensor(num_or_size_splits) else num_or_size_splits)]
    return list(torch.split(value, sizes, dim=axis))


def numpy_function(func, inp, Tout):
    return func(*inp)


# ---- training (model_trainer.py:267-273) ----------------------------------------------------------------------------
# `trace` (a list, or None) receives one record per tape.gradient / optimizer.apply_gradients call, so that
# oracle/make_golden.py can record what the reference's own training loop does around the model.
trace = None


class GradientTape:
    """``with tf.GradientTape() as tape: ...; tape.gradient(target, sources)`` over torch autograd (the shim's Variables
    are torch leaves, so everything computed inside the context is already on torch's tape)."""

    def __enter__(self):
        self._prev = torch.is_grad_enabled()
        torch.set_grad_enabled(True)
        return self

    def __exit__(self, *exc):
        torch.set_grad_enabled(self._prev)
        return False

    def gradient(self, target, sources):
        sources = list(sources)
        hook = getattr(sources[0], "_shim_gradient", None) if sources else None
        if hook is not None:  # a model that differentiates itself (the recorded-driver replay of tests/)
            grads = hook(target, sources)
        else:
            grads = list(torch.autograd.grad(target, [_val(v) for v in sources], allow_unused=True))
        if trace is not None:
            trace.append({"op": "tape_gradient", "target": float(target.detach()), "n_sources": len(sources),
                          "grads": [None if g is None else g.detach().cpu().numpy() for g in grads]})
        return grads

"""``tensorflow.keras.layers`` stand-in: Layer, Dense, Activation (test infrastructure only)."""
import numpy as np
import torch

import tensorflow as tf


class Layer:
    def __init__(self, name=None, **kwargs):
        from tensorflow.keras import _unique_name
        object.__setattr__(self, "_tracked", [])
        object.__setattr__(self, "_own_trainable", [])
        object.__setattr__(self, "_own_non_trainable", [])
        self._name = name or _unique_name(type(self).__name__)
        self.built = False
        self.trainable = True

    @property
    def name(self):
        return self._name

    def __setattr__(self, key, value):
        if isinstance(value, Layer) and value not in self._tracked:
            self._tracked.append(value)
        object.__setattr__(self, key, value)

    def add_weight(self, name, shape, trainable=True, initializer=None):
        value = initializer(tuple(shape)) if initializer is not None else torch.zeros(tuple(shape), dtype=tf._FLOAT)
        var = tf.Variable(value.to(tf._FLOAT), f"{self.name}/{name}:0", trainable)
        (self._own_trainable if trainable else self._own_non_trainable).append(var)
        return var

    @property
    def variables(self):
        return self._own_trainable + self._own_non_trainable

    @property
    def trainable_variables(self):
        return list(self._own_trainable)

    def build(self, input_shape):
        self.built = True

    def compute_output_shape(self, input_shape):
        return list(input_shape)

    def __call__(self, inputs, *args, **kwargs):
        if not self.built:
            raise RuntimeError(f"layer {self.name} called before build (the reference builds everything up front)")
        return self.call(inputs, *args, **kwargs)


def _orthogonal(shape):
    a = np.random.standard_normal((max(shape), max(shape)))
    q, r = np.linalg.qr(a)
    q = q * np.sign(np.diag(r))
    return torch.tensor(q[:shape[0], :shape[1]].copy())


class Dense(Layer):
    """y = activation(x @ kernel + bias); kernel [in, units] created in ``build``, then bias [units]."""

    def __init__(self, units, activation=None, use_bias=True, kernel_initializer=None, name=None):
        super().__init__(name=name)
        assert activation in (None, "relu") and kernel_initializer in (None, "orthogonal", "glorot_uniform")
        self.units, self.activation, self.use_bias = units, activation, use_bias
        self.kernel = self.bias = None

    def build(self, input_shape):
        if self.kernel is None:
            self.kernel = self.add_weight("kernel", (int(input_shape[-1]), self.units), True, _orthogonal)
            if self.use_bias:
                self.bias = self.add_weight("bias", (self.units,), True, lambda s: torch.zeros(s))
        self.built = True

    def compute_output_shape(self, input_shape):
        return list(input_shape[:-1]) + [self.units]

    def call(self, inputs, *args, **kwargs):
        y = inputs @ self.kernel._t
        if self.use_bias:
            y = y + self.bias._t
        return torch.relu(y) if self.activation == "relu" else y


class Activation(Layer):
    def __init__(self, activation, name=None):
        super().__init__(name=name)
        assert activation == "relu"

    def call(self, inputs, *args, **kwargs):
        return torch.relu(inputs)

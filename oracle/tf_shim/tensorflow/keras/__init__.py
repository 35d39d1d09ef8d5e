"""Torch-backed stand-in for ``tensorflow.keras`` (Model / Sequential) -- test infrastructure only.

Restates the Keras 2.7 behaviours the reference relies on:
* attribute assignment of a Layer tracks it, in first-assignment order (``Model.layers``);
* ``Model.variables`` = for each tracked layer in order: ``layer.variables``; a plain Layer lists its trainable
  weights then its non-trainable weights, each in creation order; ``trainable_variables`` likewise;
* auto-generated unique snake_case layer names (``pre_norm_layer``, ``pre_norm_layer_1`` ...).
"""
import re
from collections import defaultdict

_NAME_COUNTS = defaultdict(int)


def reset_name_counters():
    _NAME_COUNTS.clear()


def _unique_name(cls_name: str) -> str:
    snake = re.sub(r"(?<!^)(?=[A-Z][a-z])|(?<=[a-z0-9])(?=[A-Z])", "_", cls_name).lower()
    n = _NAME_COUNTS[snake]
    _NAME_COUNTS[snake] += 1
    return snake if n == 0 else f"{snake}_{n}"


from .layers import Layer  # noqa: E402


class Model(Layer):
    def __init__(self, name=None):
        super().__init__(name=name)

    @property
    def layers(self):
        return list(self._tracked)

    @property
    def variables(self):
        out = []
        for layer in self._tracked:
            out += layer.variables
        return out + self._own_trainable + self._own_non_trainable

    @property
    def trainable_variables(self):
        out = []
        for layer in self._tracked:
            out += layer.trainable_variables
        return out + self._own_trainable

    def build(self, input_shape):
        self.built = True


class Sequential(Model):
    def __init__(self, layers=None, name=None):
        super().__init__(name=name)
        for layer in layers or []:
            self._tracked.append(layer)

    def build(self, input_shape):
        shape = list(input_shape)
        for layer in self._tracked:
            layer.build(shape)
            layer.built = True
            shape = layer.compute_output_shape(shape)
        self.built = True

    def call(self, inputs, *args, **kwargs):
        for layer in self._tracked:
            inputs = layer(inputs)
        return inputs


from . import metrics  # noqa: E402,F401  (model_tester.py:199 reaches it as tf.keras.metrics)

"""``tensorflow.keras.initializers`` stand-in (test infrastructure only)."""
import torch


class Constant:
    def __init__(self, value=0.0):
        self.value = value

    def __call__(self, shape):
        return torch.full(tuple(shape), float(self.value))

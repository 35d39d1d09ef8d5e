"""``tensorflow.keras.metrics`` stand-in: mean_squared_error = mean over the last axis (model_tester.py:199 calls it on
flat [n_cuts] tensors, so the result is a scalar)."""
import torch


def mean_squared_error(y_true, y_pred):
    y_pred = torch.as_tensor(y_pred)
    y_true = torch.as_tensor(y_true).to(y_pred.dtype)
    return ((y_pred - y_true) ** 2).mean(dim=-1)

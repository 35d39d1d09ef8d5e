"""``tensorflow.keras.losses`` stand-in: MeanSquaredError on 1-D tensors = mean over all elements (Keras 2.7,
reduction AUTO -> SUM_OVER_BATCH_SIZE; model_trainer.py:271 calls it on flat [n_cuts] tensors)."""
import torch


class MeanSquaredError:
    def __call__(self, y_true, y_pred):
        y_true = torch.as_tensor(y_true)
        y_pred = torch.as_tensor(y_pred)
        return ((y_pred - y_true.to(y_pred.dtype)) ** 2).mean()

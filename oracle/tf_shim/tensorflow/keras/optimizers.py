"""``tensorflow.keras.optimizers`` stand-in -- test infrastructure only.

Adam restates Keras 2.7's dense update (optimizer_v2/adam.py, ``_resource_apply_dense`` -> ``ResourceApplyAdam``,
epsilon 1e-7, no amsgrad):  lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t);  m += (g - m)(1 - beta1);
v += (g^2 - v)(1 - beta2);  var -= lr_t * m / (sqrt(v) + eps), with t counting from 1 and ``learning_rate`` allowed to
be a callable (model_trainer.py:131 passes ``lambda: lr``).  The same arithmetic as gcnn_oracle.adam_step."""
import math

import torch

import tensorflow as tf


class Adam:
    def __init__(self, learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        self.learning_rate, self.beta_1, self.beta_2, self.epsilon = learning_rate, beta_1, beta_2, epsilon
        self.iterations = 0
        self._slots = {}

    def apply_gradients(self, grads_and_vars):
        pairs = [(g, v) for g, v in grads_and_vars if g is not None]
        lr = self.learning_rate() if callable(self.learning_rate) else self.learning_rate
        self.iterations += 1
        t = self.iterations
        if tf.trace is not None:
            tf.trace.append({"op": "apply_gradients", "lr": float(lr), "iteration": t, "n_vars": len(pairs)})
        hook = getattr(pairs[0][1], "_shim_apply", None) if pairs else None
        if hook is not None:  # a model that owns its optimiser state (the recorded-driver replay of tests/)
            hook(float(lr), t)
            return
        lr_t = lr * math.sqrt(1.0 - self.beta_2 ** t) / (1.0 - self.beta_1 ** t)
        with torch.no_grad():
            for g, var in pairs:
                m, v = self._slots.setdefault(id(var), (torch.zeros_like(var._t), torch.zeros_like(var._t)))
                g = g.to(var._t.dtype)
                m += (g - m) * (1.0 - self.beta_1)
                v += (g * g - v) * (1.0 - self.beta_2)
                var._t -= lr_t * m / (v.sqrt() + self.epsilon)

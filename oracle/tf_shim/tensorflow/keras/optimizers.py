"""``tensorflow.keras.optimizers`` stand-in: only the name is imported by model_trainer.py (the oracle restates Keras
Adam separately, gcnn_oracle.adam_step)."""


class Adam:
    def __init__(self, *args, **kwargs):
        raise NotImplementedError("the shim does not run optimisers; see gcnn_oracle.adam_step")

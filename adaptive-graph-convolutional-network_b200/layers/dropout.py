"""Dropout: models/layers/dropout.py:27-41 (active only in the train phase)."""
import torch

from ..operators import model_operatos as model_ops
from .basic_layer import Layer


class Dropout(Layer):
    def __init__(self, p, seed=None, **kwargs):
        self.p = p
        self.seed = seed
        if 0. < self.p < 1.:
            self.uses_learning_phase = True
        super(Dropout, self).__init__(**kwargs)

    def call(self, x):
        if 0. < self.p < 1.:
            def dropped_inputs():
                if self.seed is not None:
                    gen = torch.Generator(device=x.device).manual_seed(int(self.seed))
                    keep = (torch.rand(x.shape, generator=gen, device=x.device) >= self.p).to(x.dtype)
                    return x * keep / (1. - self.p)
                return torch.nn.functional.dropout(x, self.p, training=True)

            x = model_ops.in_train_phase(dropped_inputs, lambda: x)
        return x

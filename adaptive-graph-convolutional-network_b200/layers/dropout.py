"""Dropout: models/layers/dropout.py:27-41 (active only in the train phase).  The signature and the learning-phase
switch mirror the reference (category: contract-defining boundary mirror); the mask itself comes from the library's
counter-based Philox kernel (agcn_dropout, include/agcn_sgcll.h) and is regenerated from the seed in the backward
pass instead of being stored."""
import itertools

import torch

from .. import _lib
from ..batch import _ptr, _stream_ptr
from ..operators import model_operatos as model_ops
from .basic_layer import Layer

_auto_seed = itertools.count(0x5eed0001)


class _Dropout(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, p, seed):
        x = x.contiguous()
        y = torch.empty_like(x)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().agcn_dropout(_ptr(x), _ptr(y), x.numel(), float(p), int(seed), _stream_ptr(x.device)))
        ctx.p, ctx.seed = p, seed
        return y

    @staticmethod
    def backward(ctx, dy):
        dy = dy.contiguous()
        dx = torch.empty_like(dy)
        with torch.cuda.device(dy.device):
            _lib.check(_lib.lib().agcn_dropout(_ptr(dy), _ptr(dx), dy.numel(), float(ctx.p), int(ctx.seed),
                                               _stream_ptr(dy.device)))
        return dx, None, None


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
                if not x.is_cuda:
                    raise _lib.AgcnError("Dropout needs a CUDA tensor: there is no CPU path")
                # a fixed seed reproduces the mask (tf.nn.dropout(seed=...)); without one every call draws a new stream
                seed = int(self.seed) if self.seed is not None else next(_auto_seed)
                return _Dropout.apply(x, self.p, seed)

            x = model_ops.in_train_phase(dropped_inputs, lambda: x)
        return x

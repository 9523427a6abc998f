"""Activation registry: models/operators/activations.py (get :111-114, get_from_module :15-53)."""
import torch

from . import model_operatos as model_ops


def get_from_module(identifier, module_params, module_name, instantiate=False, kwargs=None):
    """activations.py:15-53: look a name up, ValueError when unknown, pass objects through."""
    if isinstance(identifier, str):
        res = module_params.get(identifier)
        if not res:
            raise ValueError('Invalid ' + str(module_name) + ': ' + str(identifier))
        if instantiate and not kwargs:
            return res()
        elif instantiate and kwargs:
            return res(**kwargs)
        return res
    return identifier


def softmax(x):
    if x.dim() in (2, 3):
        return torch.softmax(x, dim=-1)
    raise ValueError('Cannot apply softmax to a tensor that is not 2D or 3D. Here, ndim=' + str(x.dim()))


def elu(x, alpha=1.0):
    return torch.nn.functional.elu(x, alpha)


def softplus(x):
    return torch.nn.functional.softplus(x)


def softsign(x):
    return torch.nn.functional.softsign(x)


def relu(x, alpha=0., max_value=None):
    return model_ops.relu(x, alpha=alpha, max_value=max_value)


def tanh(x):
    return torch.tanh(x)


def sigmoid(x):
    return torch.sigmoid(x)


def hard_sigmoid(x):
    return torch.clamp(0.2 * x + 0.5, 0., 1.)


def linear(x):
    return x


def get(identifier):
    """activations.py:111-114."""
    if identifier is None:
        return linear
    return get_from_module(identifier, globals(), 'activation function')

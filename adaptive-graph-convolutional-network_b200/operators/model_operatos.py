"""The slice of models/operators/model_operatos.py that sits on the SGC-LL path: the leaky
rectifier applied to the Laplacian and the layer outputs (model_operatos.py:534-558), the
learning-phase switch used by Dropout (:32-95) and the layer-name counter (:156-168)."""
import collections

import torch

_UID_PREFIXES = collections.defaultdict(int)
_LEARNING_PHASE = [1]  # 1 = train, 0 = test (the reference feeds a bool placeholder)


def get_uid(prefix=''):
    """model_operatos.py:156-168."""
    _UID_PREFIXES[prefix] += 1
    return _UID_PREFIXES[prefix]


def reset_uids():
    _UID_PREFIXES.clear()


def learning_phase():
    """model_operatos.py:32-43: 0 = test, 1 = train."""
    return _LEARNING_PHASE[0]


def set_learning_phase(value):
    if value not in (0, 1):
        raise ValueError('Expected learning phase to be 0 or 1.')
    _LEARNING_PHASE[0] = value


def in_train_phase(x, alt):
    """model_operatos.py:46-61: `x` in the train phase, `alt` otherwise (both may be callables)."""
    chosen = x if learning_phase() == 1 else alt
    return chosen() if callable(chosen) else chosen


def relu(x, alpha=0., max_value=None):
    """model_operatos.py:534-558: relu(x) - alpha * relu(-x), optional saturation.  `alpha` may be
    a float or a (1,)-tensor variable (a Variable is always `!= 0.` in the reference, :548)."""
    use_alpha = isinstance(alpha, torch.Tensor) or alpha != 0.
    if use_alpha:
        negative_part = torch.relu(-x)
    x = torch.relu(x)
    if max_value is not None:
        x = torch.clamp(x, 0., max_value)
    if use_alpha:
        x = x - alpha * negative_part
    return x

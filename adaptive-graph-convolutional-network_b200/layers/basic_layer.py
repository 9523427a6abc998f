"""Layer base class: models/layers/basic_layer.py (kwargs check :57-62, auto-naming :63-67,
__call__ -> call :155-161)."""
from ..operators import model_operatos as model_ops


class Layer(object):
    def __init__(self, **kwargs):
        if not hasattr(self, 'uses_learning_phase'):
            self.uses_learning_phase = False
        if not hasattr(self, 'losses'):
            self.losses = []
        allowed_kwargs = {'input_shape', 'batch_input_shape', 'input_dtype', 'name', 'trainable'}
        for kwarg in kwargs.keys():
            if kwarg not in allowed_kwargs:
                raise TypeError('Keyword argument not understood:', kwarg)
        name = kwargs.get('name')
        if not name:
            prefix = self.__class__.__name__.lower()
            name = prefix + '_' + str(model_ops.get_uid(prefix))
        self.name = name
        self.trainable = kwargs.get('trainable', True)
        if 'batch_input_shape' in kwargs or 'input_shape' in kwargs:
            if 'batch_input_shape' in kwargs:
                batch_input_shape = tuple(kwargs['batch_input_shape'])
            else:
                batch_input_shape = (None,) + tuple(kwargs['input_shape'])
            self.batch_input_shape = batch_input_shape
            self.input_dtype = kwargs.get('input_dtype', 'float64')

    def call(self, x):
        return x

    def __call__(self, x):
        return self.call(x)

    @staticmethod
    def to_list(x):
        if isinstance(x, list):
            return x
        return [x]

    def parameters(self):
        """Trainable tensors of the layer (the reference keeps them in self.vars)."""
        return [v for v in getattr(self, 'vars', {}).values() if getattr(v, 'requires_grad', False)]

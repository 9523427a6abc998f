from . import activations
from . import model_operatos
from .model_operatos import relu, in_train_phase, learning_phase, set_learning_phase, get_uid

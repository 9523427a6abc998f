"""models/layers/__init__.py:16-29 -- the classes on the SGC-LL path."""
from .basic_layer import Layer
from .dropout import Dropout
from .graphconv import SGC_LL, glorot, zeros, truncate_normal
from .graphconv_reslap import SGC_LL_Reslap
from .graphpool import GraphPoolMol
from .blockend import BlockEnd
from .densenet_block import DenseBlockEnd
from .MLP import MLP

"""agcn_b200 -- B200-native SGC-LL hot path of uta-smile/Adaptive-Graph-Convolutional-Network.

The directory is named ``adaptive-graph-convolutional-network_b200``; import it as ``agcn_b200``
(the shim package at the repository root).  Mirrors the reference's ``models`` package for the hot
path only: ``Graph`` / ``MolGraph`` (models/graph_structure.py), ``layers.SGC_LL`` /
``layers.SGC_LL_Reslap`` (models/layers), ``operators`` (models/operators).
"""
from . import _lib
from .batch import GraphBatch, PackedNodes, PackedLaplacians
from .graph_structure import Graph, MolGraph
from .graph_topology import GraphTopologyMol
from . import layers
from . import operators
from .layers import SGC_LL, SGC_LL_Reslap, GraphPoolMol, BlockEnd, DenseBlockEnd, MLP

__all__ = ["Graph", "MolGraph", "GraphTopologyMol", "GraphBatch", "PackedNodes", "PackedLaplacians", "layers",
           "operators", "SGC_LL", "SGC_LL_Reslap", "GraphPoolMol", "BlockEnd", "DenseBlockEnd", "MLP"]

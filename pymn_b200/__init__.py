"""pymn_b200 -- B200-native MetaNeighbor scoring hot path behind pyMN's API.

Drop-in for the three hot-path entry points of gillislab/pyMN
(reference pymn/__init__.py:2-4,18):

    from pymn_b200 import MetaNeighborUS, MetaNeighbor, trainModel, join_labels

Importing the package does not need a GPU; calling any entry point does, and
raises if the CUDA library (``libpymn_b200.so``) or a CUDA device is missing --
there is no CPU fallback.
"""
from .MetaNeighbor import MetaNeighbor
from .MetaNeighborUS import MetaNeighborUS
from .trainModel import trainModel
from .utils import join_labels
from .anndata_lite import AnnDataLite

__all__ = ["MetaNeighborUS", "MetaNeighbor", "trainModel", "join_labels", "AnnDataLite"]
__version__ = "0.1.0"

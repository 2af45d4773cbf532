"""pymn_b200 -- B200-native MetaNeighbor scoring hot path behind pyMN's API.

Drop-in for the hot-path entry points of gillislab/pyMN and the step right before them
(reference pymn/__init__.py:2-5,18):

    from pymn_b200 import MetaNeighborUS, MetaNeighbor, trainModel, variableGenes, join_labels

Importing the package does not need a GPU; calling any entry point does, and
raises if the CUDA library (``libpymn_b200.so``) or a CUDA device is missing --
there is no CPU fallback.
"""
from .MetaNeighbor import MetaNeighbor
from .MetaNeighborUS import MetaNeighborUS
from .trainModel import trainModel
from .variableGenes import variableGenes
from .utils import join_labels
from .anndata_lite import AnnDataLite
from .model_io import load_model, save_model



def release_caches():
    """Give back the device memory the package keeps between calls: the rho spill of the full-network path (up to
    ~80 GB for 200,000 cells: allocating it costs more than a call, so it is cached -- ``MN_SPILL_GB`` caps it, ``0``
    disables the spill), the multi-GPU peer receive buffers, the device buffer of the streamed upload and the pinned staging area.  On several GPUs call
    it on every rank."""
    from . import engine
    engine.release_spill_cache()
    engine.release_peer_buffers()
    engine._STAGE.clear()
    engine._XD_CACHE.clear()


__all__ = ["MetaNeighborUS", "MetaNeighbor", "trainModel", "variableGenes", "join_labels", "AnnDataLite", "save_model", "load_model",
           "release_caches"]
__version__ = "0.1.0"

"""Load the UNMODIFIED reference hot-path modules by file path -- BUILD-CONTAINER ONLY.

TEST INFRASTRUCTURE.  ``/root/reference`` exists only in the build container,
never on the GPU box, so this loader is used by ``oracle/make_golden.py`` (to
produce the committed fixtures under ``tests/golden/``) and by the
``needs_reference`` tests that cross-check the restatement
(``oracle/mn_oracle.py``) against the reference live.  Nothing at run time on
the GPU box may import it.

``import pymn`` itself is impossible here: ``pymn/__init__.py:5-14`` pulls in
plotting (seaborn, scanpy, upsetplot, matplotlib) and every hot-path module
imports the C extension ``bottleneck``; none are installed and there is no
network.  The recipe: register a synthetic ``pymn``
package, shim ``bottleneck``'s five functions with their SciPy / NumPy
equivalents (average-tie, 1-based, float64 -- the documented semantics of
``bottleneck.rankdata`` / ``nanrankdata``), and exec
``pymn/{utils,MetaNeighborUS,MetaNeighbor,trainModel,variableGenes}.py`` as they lie.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types
import warnings

import numpy as np

REFERENCE_ROOT = os.environ.get("PYMN_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "pymn", "utils.py"))


def _bottleneck_shim():
    import scipy.stats
    bn = types.ModuleType("bottleneck")

    def rankdata(a, axis=None):
        a = np.asarray(a)
        if axis is None:
            return scipy.stats.rankdata(a.ravel(), "average").astype(np.float64)
        return scipy.stats.rankdata(a, "average", axis=axis).astype(np.float64)

    def nanrankdata(a, axis=None):
        return scipy.stats.rankdata(np.asarray(a, float), "average", axis=axis, nan_policy="omit")

    bn.rankdata = rankdata
    bn.nanrankdata = nanrankdata
    bn.nansum = np.nansum
    bn.nanmean = np.nanmean
    bn.median = np.median
    return bn


_loaded = None


def load_reference():
    """Returns a namespace with the reference modules: .utils, .MetaNeighborUS,
    .MetaNeighbor, .trainModel, .variableGenes (module objects)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    import pandas as pd
    # pandas-3 default string dtype breaks utils.py:119 (design_matrix -> set_index
    # on .values of a str column); the reference predates it.
    try:
        pd.set_option("future.infer_string", False)
    except Exception:
        pass
    sys.modules.setdefault("bottleneck", _bottleneck_shim())
    pkg = types.ModuleType("pymn")
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, "pymn")]
    sys.modules["pymn"] = pkg
    ns = types.SimpleNamespace()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")          # SyntaxWarning: `is not "highly_variable"` (trainModel.py:29)
        for name in ("utils", "MetaNeighborUS", "MetaNeighbor", "trainModel", "variableGenes"):
            path = os.path.join(REFERENCE_ROOT, "pymn", f"{name}.py")
            spec = importlib.util.spec_from_file_location(f"pymn.{name}", path)
            mod = importlib.util.module_from_spec(spec)
            sys.modules[f"pymn.{name}"] = mod
            spec.loader.exec_module(mod)
            setattr(ns, name, mod)
    _loaded = ns
    return ns

"""pmf_b200 — B200-native (sm_100a) implementation of the ICEORY/PMF data-parallel hot path.

Public surface (mirrors the reference's names):
    pmf_b200.PMFNet, pmf_b200.ResidualBasedFusionBlock      pc_processor/models/pmf_net.py
    pmf_b200.EPMFNet (inference only)                       pc_processor/models/epmf_net.py
    pmf_b200.KNN                                            pc_processor/postproc/knn.py
    pmf_b200.project_scatter                                perspective projection (parser.py:209-227 + loader scatter)
Everything computes through libpmf_b200.so (include/pmfb.h); there is no CPU path.
"""
from .modules import EPMFNet, PMFNet, ResidualBasedFusionBlock  # noqa: F401
from .postproc import KNN, InferenceTail, argmax_nchw, knn_batched, lut_remap, merge_cameras, project_scatter  # noqa: F401


import contextlib as _contextlib

from ._lib import get_precision, set_precision  # noqa: E402,F401


@_contextlib.contextmanager
def precision(mode):
    """``with pmf_b200.precision("3xtf32"): ...`` — run the tensor-core convolutions in the precise mode (see
    pmf_b200/_lib.py; also selectable process-wide with the environment variable PMFB_PRECISION)."""
    prev = set_precision(mode)
    try:
        yield
    finally:
        set_precision(prev)

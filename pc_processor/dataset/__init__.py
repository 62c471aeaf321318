"""pc_processor.dataset — the reference's names (pc_processor/dataset/__init__.py:1-8).  ``PerspectiveViewLoader`` is the
B200 version (projection + scatter on the device, pmf_b200/loader.py); the dataset parsers and the other loaders are
host-side file I/O outside the accelerated path and are imported from the reference's own files when a deployment keeps
them next to this one."""
import importlib as _importlib
import importlib.util as _util

from .perspective_view_loader import PerspectiveViewLoader  # noqa: F401


def _have(name):
    return _util.find_spec(__name__ + "." + name) is not None


for _sub in ("semantic_kitti", "nuScenes", "a2d2"):
    if _have(_sub):
        _importlib.import_module(__name__ + "." + _sub)
if _have("salsanext_loader"):
    from .salsanext_loader import SalsaNextLoader  # noqa: F401
if _have("sensat_urban"):
    from .sensat_urban import SensatLoader, SensatUrban  # noqa: F401
if _have("perspective_view_loader_v2"):
    from .perspective_view_loader_v2 import PerspectiveViewLoaderV2  # noqa: F401
del _sub

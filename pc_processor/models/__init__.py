"""pc_processor.models — the names the reference exports (pc_processor/models/__init__.py:1-3).  PMFNet and EPMFNet are
the B200 implementations; SalsaNext (the LiDAR-only baseline of tasks/salsanext, outside the accelerated path) is
re-exported from the reference's own ``salsanext.py`` when a deployment keeps that file next to this one
(INTEGRATION.md §1)."""
import importlib.util as _util

from .pmf_net import PMFNet  # noqa: F401
from .epmf_net import EPMFNet  # noqa: F401

if _util.find_spec(__name__ + ".salsanext") is not None:
    from .salsanext import SalsaNext  # noqa: F401

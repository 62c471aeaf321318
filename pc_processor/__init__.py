"""Drop-in import surface for the reference's task scripts.  ``tasks/pmf/main.py`` does a bare
``import pc_processor`` and then reaches ``pc_processor.utils.init_distributed_mode``, ``.checkpoint.Recorder``,
``.models.PMFNet``, ``.layers.sync_bn.replaceBN``, ``.loss``, ``.metrics`` and ``.dataset`` by attribute
(/root/reference/pc_processor/__init__.py:1-8 imports all eight sub-packages), so this package does the same.

The hot-path sub-packages live here and forward to pmf_b200 (libpmf_b200.so): ``models``, ``postproc``, and the
GPU-projection ``dataset/perspective_view_loader.py`` override.  The remaining sub-packages (checkpoint, layers, loss,
metrics, utils and the rest of dataset) are plain PyTorch/numpy host code outside the accelerated path and are NOT
re-implemented: a deployment keeps the reference's own copies next to this package
(INTEGRATION.md §1); each one is imported when it is present, exactly as the reference ``__init__`` would.
"""
import importlib as _importlib
import importlib.util as _util

from . import models, postproc  # noqa: F401

for _sub in ("checkpoint", "dataset", "layers", "loss", "metrics", "utils"):
    # absent sub-package (this repository on its own): skip.  Present but failing to import (e.g. tensorboardX or the
    # nuscenes devkit missing): raise, like the reference package does.
    if _util.find_spec(__name__ + "." + _sub) is not None:
        _importlib.import_module(__name__ + "." + _sub)
del _sub

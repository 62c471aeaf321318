"""Drop-in import surface for the reference's task scripts: ``tasks/pmf/main.py`` does
``pc_processor.models.PMFNet(...)`` and ``tasks/pmf_eval_semantickitti/infer.py`` does ``pc_processor.postproc.KNN(...)``.
Only the hot-path sub-packages exist here (models, postproc); they forward to pmf_b200 (libpmf_b200.so).  The
reference's remaining sub-packages (dataset, loss, metrics, layers, utils, checkpoint) are plain PyTorch/numpy host code
outside the accelerated path and are intentionally NOT re-implemented: a deployment keeps the
reference's own copies of those next to this package.
"""
from . import models, postproc  # noqa: F401

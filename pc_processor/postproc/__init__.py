"""pc_processor.postproc — KNN back-projection (pc_processor/postproc/knn.py:38-143)."""
from .knn import KNN  # noqa: F401

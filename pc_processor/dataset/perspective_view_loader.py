"""pc_processor.dataset.perspective_view_loader — the reference module path
(pc_processor/dataset/perspective_view_loader.py:8-141), served by the device-projection loader."""
from pmf_b200.loader import PerspectiveViewLoader  # noqa: F401

"""pc_processor.models — the names the reference exports (pc_processor/models/__init__.py:1-3) that are on the hot path."""
from .pmf_net import PMFNet  # noqa: F401
from .epmf_net import EPMFNet  # noqa: F401

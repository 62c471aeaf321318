"""pc_processor.models.pmf_net — same module path and class names as the reference file (pmf_net.py:10-36, 224-249)."""
from pmf_b200.modules import ASPP, PMFNet, ResidualBasedFusionBlock, ResNet, RGBDecoder, SalsaNextFusion  # noqa: F401

"""pc_processor.models.epmf_net — the reference module path of EPMFNet (pc_processor/models/epmf_net.py:185-216),
served by the B200 implementation (inference only in this round)."""
from pmf_b200.modules import EPMFNet, ResidualBasedFusionBlock, SparseVariantConv  # noqa: F401

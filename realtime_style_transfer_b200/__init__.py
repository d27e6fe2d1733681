"""B200-native stylization hot path behind the reference's Python surface.

    from realtime_style_transfer_b200.shape_config import ShapeConfig
    from realtime_style_transfer_b200.models import styleTransfer, stylePrediction, styleTransferInferenceModel

mirrors ``realtime_style_transfer`` of singinwhale/realtime-style-transfer for the path named in
SURVEY.md section 8.  All arithmetic runs in csrc/librst_sm100.so (hand-written sm_100a CUDA).
"""
from . import mixed_precision  # noqa: F401
from .shape_config import ShapeConfig  # noqa: F401

__version__ = "0.1.0"

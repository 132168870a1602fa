"""kgnet-b200: B200-native KGnet inference hot path behind the reference's Python call surface.

    from kg_instance_segmentation_b200 import KGnet, postprocessing, nms

All compute runs in hand-written sm_100a CUDA kernels (csrc/) behind a C-ABI shared library
(include/kgnet_b200.h); there is no CPU or PyTorch-op fallback.
"""
from . import config  # noqa: F401
from . import _cabi  # noqa: F401
from . import nms, postprocessing  # noqa: F401
from . import KGnet  # noqa: F401
from . import preprocessing, loss, seg_loss  # noqa: F401

__all__ = ["config", "nms", "postprocessing", "KGnet", "preprocessing", "loss", "seg_loss"]

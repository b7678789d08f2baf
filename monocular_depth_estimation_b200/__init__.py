"""B200-native (sm_100a) NeWCRFs CRF-block hot path: drop-in modules over a hand-written CUDA C-ABI library.

    from monocular_depth_estimation_b200 import BasicCRFLayer, CRFBlock, WindowAttention, NewCRF

See include/crf_sm100.h for the C ABI, DESIGN.md for the kernel design and INTEGRATION.md for how the reference
repository adopts it.
"""
from .newcrf_layers import BasicCRFLayer, CRFBlock, Mlp, NewCRF, WindowAttention  # noqa: F401
from .functional import crf_block, convert_v  # noqa: F401
from .ops import set_precision  # noqa: F401

__all__ = ["BasicCRFLayer", "CRFBlock", "Mlp", "NewCRF", "WindowAttention", "crf_block", "convert_v", "set_precision"]

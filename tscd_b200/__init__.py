"""tscd_b200 -- B200-native (sm_100a) video-level proposal aggregation stage of TSCD.

Python host code + hand-written CUDA kernels behind a C-ABI (include/tscd_b200.h).  The package fails
loudly when libtscd_b200.so is missing; there is no CPU / PyTorch fallback on the product path.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
__version__ = "0.1"

"""pyremap_b200: pyremap's weight-application hot path on NVIDIA B200 (sm_100a).

Drop-in for ``pyremap.Remapper(...).remap_numpy(ds, renormalization_threshold)``
(alias ``remap``); see DESIGN.md.  Hand-written CUDA behind a C ABI
(``include/b200remap.h``); PyTorch only carries device buffers.
"""

from .remapper import Remapper  # noqa: F401
from ._cabi import B200RemapError  # noqa: F401

__version__ = '0.1.0'

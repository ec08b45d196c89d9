"""spine_vision_b200 -- B200-native (sm_100a) localization-and-crop hot path of spine-vision.

Host side: Python/PyTorch plumbing that mirrors the reference's callables
(``spine_vision_b200.cropping``) plus the batched driver (``pipeline``).
Device side: hand-written CUDA behind the C ABI in ``include/spine_b200.h``
(``libspine_b200.so``, built by ``__graft_entry__.build()``).  No CPU fallback.
"""

__version__ = "0.1.0"

from . import _lib  # noqa: F401


def library_path():
    return _lib.LIB_PATH

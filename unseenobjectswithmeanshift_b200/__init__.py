"""B200-native (sm_100a) implementation of the MSMFormer segmentation hot path.

Drop-in for the hot path of YoungSean/UnseenObjectsWithMeanShift: the
``meanshiftformer`` sub-package mirrors the reference's module/class/registry names and
``state_dict`` layout; every hot op runs in hand-written CUDA behind the C ABI declared in
``include/msmformer_b200.h`` (``lib/libmsmformer_b200.so``, sources in ``csrc/``).

There is no CPU fallback: using an op without the built library or without a CUDA tensor raises.
"""
from . import _lib  # noqa: F401  (does not load the library until first use)

__version__ = "0.1.0"

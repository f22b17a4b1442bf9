from . import modeling  # noqa: F401  (registers heads / pixel decoders / transformer decoders)
from .meanshiftformer_model import MeanShiftMaskFormer, PretrainedMeanShiftMaskFormer  # noqa: F401,E402  (META_ARCH registry)

"""Reference module name (MSMFormer/meanshiftformer/pretrained_meanshiftformer_model.py): re-exports the wrapper."""
from .meanshiftformer_model import PretrainedMeanShiftMaskFormer  # noqa: F401

"""Importing this package performs the registrations the reference's modeling/__init__.py does."""
from .pixel_decoder.fpn import SimpleBasePixelDecoder, build_pixel_decoder  # noqa: F401
from .pixel_decoder.msdeformattn import MSDeformAttnPixelDecoder  # noqa: F401
from .meta_arch.meanshift_former_head import MeanShiftMaskFormerHead, PretrainedMeanShiftMaskFormerHead  # noqa: F401
from .transformer_decoder.meanshiftformer_transformer_decoder import (  # noqa: F401
    MeanShiftTransformerDecoder, PretrainedMeanShiftTransformerDecoder)

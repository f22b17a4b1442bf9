"""Registry of transformer decoders (reference modeling/transformer_decoder/maskformer_transformer_decoder.py:15-27)."""
from ....d2compat import Registry

TRANSFORMER_DECODER_REGISTRY = Registry("TRANSFORMER_MODULE")
TRANSFORMER_DECODER_REGISTRY.__doc__ = "Registry for transformer module in MaskFormer."


def build_transformer_decoder(cfg, in_channels, mask_classification=True):
    name = cfg.MODEL.MASK_FORMER.TRANSFORMER_DECODER_NAME
    return TRANSFORMER_DECODER_REGISTRY.get(name)(cfg, in_channels, mask_classification)

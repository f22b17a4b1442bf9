"""Segmentation head glue - mirror of the reference's modeling/meta_arch/meanshift_former_head.py:145-275
(``PretrainedMeanShiftMaskFormerHead``) and :18-143 (``MeanShiftMaskFormerHead``, same forward)."""
from typing import Dict

from torch import nn

from ....d2compat import SEM_SEG_HEADS_REGISTRY, ShapeSpec, configurable
from ..pixel_decoder.fpn import build_pixel_decoder
from ..transformer_decoder.maskformer_transformer_decoder import build_transformer_decoder


class _HeadBase(nn.Module):
    _version = 2

    @configurable
    def __init__(self, input_shape: Dict[str, ShapeSpec], *, num_classes: int, pixel_decoder: nn.Module,
                 loss_weight: float = 1.0, ignore_value: int = -1, transformer_predictor: nn.Module,
                 transformer_in_feature: str):
        super().__init__()
        by_stride = sorted(input_shape.items(), key=lambda kv: kv[1].stride)
        self.in_features = [k for k, _ in by_stride]
        self.ignore_value = ignore_value
        self.common_stride = 4
        self.loss_weight = loss_weight
        self.pixel_decoder = pixel_decoder
        self.predictor = transformer_predictor
        self.transformer_in_feature = transformer_in_feature
        self.num_classes = num_classes

    @classmethod
    def from_config(cls, cfg, input_shape: Dict[str, ShapeSpec]):
        mf, head = cfg.MODEL.MASK_FORMER, cfg.MODEL.SEM_SEG_HEAD
        if mf.TRANSFORMER_IN_FEATURE in ("transformer_encoder", "multi_scale_pixel_decoder"):
            in_channels = head.CONVS_DIM
        elif mf.TRANSFORMER_IN_FEATURE == "pixel_embedding":
            in_channels = head.MASK_DIM
        else:
            in_channels = input_shape[mf.TRANSFORMER_IN_FEATURE].channels
        return {
            "input_shape": {k: v for k, v in input_shape.items() if k in head.IN_FEATURES},
            "ignore_value": head.IGNORE_VALUE,
            "num_classes": head.NUM_CLASSES,
            "pixel_decoder": build_pixel_decoder(cfg, input_shape),
            "loss_weight": head.LOSS_WEIGHT,
            "transformer_in_feature": mf.TRANSFORMER_IN_FEATURE,
            "transformer_predictor": build_transformer_decoder(cfg, in_channels, mask_classification=True),
        }

    def forward(self, features, image_height, image_width, mask=None):
        return self.layers(features, image_height, image_width, mask)

    def layers(self, features, image_height, image_width, mask=None):
        mask_features, _, multi_scale_features = self.pixel_decoder.forward_features(features)
        if self.transformer_in_feature != "multi_scale_pixel_decoder":
            # the reference calls exit() here (:273-274); raise instead of killing the process
            raise NotImplementedError("only TRANSFORMER_IN_FEATURE == 'multi_scale_pixel_decoder' is supported")
        predictions = self.predictor(multi_scale_features, mask_features, mask)
        return predictions, mask_features


@SEM_SEG_HEADS_REGISTRY.register()
class PretrainedMeanShiftMaskFormerHead(_HeadBase):
    pass


@SEM_SEG_HEADS_REGISTRY.register()
class MeanShiftMaskFormerHead(_HeadBase):
    pass

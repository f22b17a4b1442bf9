"""META_ARCH wrappers - mirror of the reference's ``MeanShiftMaskFormer`` (meanshiftformer_model.py:38-450) and
``PretrainedMeanShiftMaskFormer`` (pretrained_meanshiftformer_model.py:47-497): same registry names, constructor
keywords, sub-module names (``backbone`` / ``pretrained_backbone``, ``sem_seg_head``, ``criterion``) and therefore the
same ``state_dict`` layout, and the same ``forward(batched_inputs) -> [{"instances": ...}]`` contract in eval mode.

Scope: the backbone is whatever module the caller supplies (ResNet-50 / UCN are cuDNN work outside the
hot path); the head runs the CUDA path of this package; the eval tail is ``instance_inference.inference_tail`` (top-k
first, one fused pass). In training mode ``forward`` returns the weighted loss dict of the reference's train branch
(pretrained_meanshiftformer_model.py:303-334: prepare_targets, SetCriterion, weight_dict scaling - row f4); the UCN
embedding loss (``use_embedding_loss``, lib/fcn code that trains the backbone) and the semantic / panoptic outputs
(unused by every UOIS config) raise ``NotImplementedError``.
"""
from typing import Tuple

import torch
from torch import nn
from torch.nn import functional as F

from ..d2compat import META_ARCH_REGISTRY, configurable
from . import instance_inference as _tail
from .modeling.criterion import SetCriterion
from .modeling.matcher import HungarianMatcher


def build_criterion(num_classes, *, class_weight=1.0, mask_weight=20.0, dice_weight=1.0, no_object_weight=0.1,
                    deep_supervision=True, dec_layers=10, train_num_points=112 * 112, oversample_ratio=3.0,
                    importance_sample_ratio=0.75):
    """The criterion the reference's from_config assembles (pretrained_meanshiftformer_model.py:165-201); defaults are
    the values of meanshiftformer/config.py:31-35,129-135 (every UOIS YAML keeps them)."""
    matcher = HungarianMatcher(cost_class=class_weight, cost_mask=mask_weight, cost_dice=dice_weight,
                               num_points=train_num_points)
    weight_dict = {"loss_ce": class_weight, "loss_mask": mask_weight, "loss_dice": dice_weight}
    if deep_supervision:
        aux = {}
        for i in range(dec_layers - 1):
            aux.update({k + f"_{i}": v for k, v in weight_dict.items()})
        weight_dict.update(aux)
    return SetCriterion(num_classes, matcher=matcher, weight_dict=weight_dict, eos_coef=no_object_weight,
                        losses=["labels", "masks"], num_points=train_num_points, oversample_ratio=oversample_ratio,
                        importance_sample_ratio=importance_sample_ratio)


class _CriterionState(nn.Module):
    """Keeps ``criterion.empty_weight`` (modeling/criterion.py:113-115) in the state_dict when no criterion is
    supplied, so reference checkpoints load with strict=True."""

    def __init__(self, num_classes, eos_coef=0.1):
        super().__init__()
        w = torch.ones(num_classes + 1)
        w[-1] = eos_coef
        self.register_buffer("empty_weight", w)


def _batch_images(images, size_divisibility):
    """ImageList.from_tensors: pad bottom/right with zeros to the largest image, rounded up to size_divisibility.
    Returns (tensor [B,C,H,W], [(h, w)] per image)."""
    if isinstance(images, torch.Tensor):
        images = list(images)
    sizes = [(int(t.shape[-2]), int(t.shape[-1])) for t in images]
    H, W = max(s[0] for s in sizes), max(s[1] for s in sizes)
    if size_divisibility and size_divisibility > 1:
        d = int(size_divisibility)
        H, W = (H + d - 1) // d * d, (W + d - 1) // d * d
    if all(s == (H, W) for s in sizes):
        return torch.stack(images), sizes
    out = images[0].new_zeros(len(images), images[0].shape[0], H, W)
    for i, t in enumerate(images):
        out[i, :, :t.shape[-2], :t.shape[-1]] = t
    return out, sizes


class _MetaArchBase(nn.Module):
    def _common(self, sem_seg_head, criterion, num_queries, object_mask_threshold, overlap_threshold, metadata,
                size_divisibility, sem_seg_postprocess_before_inference, pixel_mean, pixel_std, semantic_on,
                panoptic_on, instance_on, test_topk_per_image, use_embedding_loss):
        self.sem_seg_head = sem_seg_head
        # the eval branch never reads aux_outputs (reference :335-378): let the decoder skip their full-resolution masks
        predictor = getattr(sem_seg_head, "predictor", None)
        if predictor is not None and hasattr(predictor, "eval_aux_masks"):
            predictor.eval_aux_masks = False
        self.criterion = criterion if criterion is not None else _CriterionState(sem_seg_head.num_classes)
        self.num_queries = num_queries
        self.overlap_threshold = overlap_threshold
        self.object_mask_threshold = object_mask_threshold
        self.metadata = metadata
        self.size_divisibility = size_divisibility
        self.sem_seg_postprocess_before_inference = sem_seg_postprocess_before_inference
        self.register_buffer("pixel_mean", torch.Tensor(pixel_mean).view(-1, 1, 1), False)
        self.register_buffer("pixel_std", torch.Tensor(pixel_std).view(-1, 1, 1), False)
        self.semantic_on, self.instance_on, self.panoptic_on = semantic_on, instance_on, panoptic_on
        self.test_topk_per_image = test_topk_per_image
        self.use_embedding_loss = use_embedding_loss
        if not semantic_on:
            assert sem_seg_postprocess_before_inference

    @property
    def device(self):
        return self.pixel_mean.device

    def _check_mode(self):
        if self.training:
            raise RuntimeError("inference helper called in training mode: call .eval() first")

    def forward(self, batched_inputs):
        """Reference forward: the weighted loss dict in training mode (:303-334), [{"instances": ...}] per image in
        eval mode (:335-378)."""
        if self.training:
            return self._losses(batched_inputs)
        return self._eval_tail(*self._head_outputs(batched_inputs))

    def prepare_targets(self, targets, padded_size):
        """Reference :380-395: ground-truth masks zero-padded to the padded batch size. ``targets``: per image an
        object with ``gt_masks`` [T,h,w] and ``gt_classes`` [T] (detectron2 Instances) or a dict with those keys."""
        h_pad, w_pad = padded_size
        out = []
        for t in targets:
            get = (lambda k: t[k]) if isinstance(t, dict) else (lambda k: getattr(t, k))
            gt_masks = get("gt_masks")
            gt_masks = getattr(gt_masks, "tensor", gt_masks).to(self.device)  # BitMasks or a plain tensor
            padded = torch.zeros((gt_masks.shape[0], h_pad, w_pad), dtype=gt_masks.dtype, device=self.device)
            padded[:, :gt_masks.shape[1], :gt_masks.shape[2]] = gt_masks
            out.append({"labels": get("gt_classes").to(self.device), "masks": padded})
        return out

    def _losses(self, batched_inputs, point_source=None):
        if self.use_embedding_loss:
            raise NotImplementedError("the UCN embedding loss (lib/fcn/... EmbeddingLoss) trains the embedding "
                                      "backbone, which is outside this package (SURVEY.md 8, out of scope)")
        if not isinstance(self.criterion, SetCriterion):
            raise RuntimeError("training needs a criterion: pass criterion=build_criterion(num_classes, ...)")
        if "instances" not in batched_inputs[0]:
            raise ValueError('training inputs need "instances" (gt_masks, gt_classes) per image')
        outputs, _, padded_size, _ = self._head_outputs(batched_inputs)
        targets = self.prepare_targets([x["instances"] for x in batched_inputs], padded_size)
        losses = self.criterion(outputs, targets, point_source)
        weights = self.criterion.weight_dict
        return {k: v * weights[k] for k, v in losses.items() if k in weights}  # others are dropped (:329-334)

    def label_maps(self, batched_inputs, topk=False, score=0.7, low_threshold=0.4):
        """What the UOIS test scripts do with the instances (lib/fcn/test_utils.py:35-52, 93-112, 216-242:
        get_confident_instances + combine_masks), fused and batched: -> (label map fp32 [B,H,W] on the device,
        per-instance fields). The full-resolution masks are never written."""
        from ..fcn import test_utils as tu
        self._check_mode()
        outputs, _, padded_size, sizes = self._head_outputs(batched_inputs)
        if any(tuple(sz) != tuple(padded_size) for sz in sizes):
            raise NotImplementedError("images of different sizes in one batch: crop the label maps yourself")
        return tu.label_map_from_outputs(outputs["pred_logits"], outputs["pred_masks"], padded_size,
                                         self.test_topk_per_image, topk=topk, score=score,
                                         num_class=self.sem_seg_head.num_classes, low_threshold=low_threshold)

    def _eval_tail(self, outputs, batched_inputs, padded_size, image_sizes):
        if self.semantic_on or self.panoptic_on or not self.instance_on:
            raise NotImplementedError("only the instance output (MODEL.MASK_FORMER.TEST.INSTANCE_ON) is implemented")
        for inp, size in zip(batched_inputs, image_sizes):
            want = (inp.get("height", size[0]), inp.get("width", size[1])) if isinstance(inp, dict) else size
            if tuple(size) != tuple(padded_size) or tuple(want) != tuple(padded_size):
                raise NotImplementedError(
                    f"output size {want} / image size {size} differ from the padded batch size {padded_size}: the "
                    "resize of sem_seg_postprocess is not implemented (the UOIS scripts run at the input size)")
        return _tail.inference_tail(outputs, padded_size, self.test_topk_per_image)


@META_ARCH_REGISTRY.register()
class MeanShiftMaskFormer(_MetaArchBase):
    """Reference meanshiftformer_model.py:38-450 (standard detectron2 backbone, mean / std normalisation :244)."""

    @configurable
    def __init__(self, *, backbone: nn.Module, sem_seg_head: nn.Module, criterion: nn.Module, num_queries: int,
                 object_mask_threshold: float, overlap_threshold: float, metadata, size_divisibility: int,
                 sem_seg_postprocess_before_inference: bool, pixel_mean: Tuple[float], pixel_std: Tuple[float],
                 semantic_on: bool, panoptic_on: bool, instance_on: bool, test_topk_per_image: int,
                 use_embedding_loss: bool = False, embedding_loss_weight: float = 1.0, alpha: float = 0.02,
                 delta: float = 0.5, lambda_intra: float = 1.0, lambda_inter: float = 1.0, metric: str = "cosine",
                 normalize: bool = True):
        super().__init__()
        self.backbone = backbone
        if size_divisibility < 0:
            size_divisibility = self.backbone.size_divisibility
        self._common(sem_seg_head, criterion, num_queries, object_mask_threshold, overlap_threshold, metadata,
                     size_divisibility, sem_seg_postprocess_before_inference, pixel_mean, pixel_std, semantic_on,
                     panoptic_on, instance_on, test_topk_per_image, use_embedding_loss)

    @classmethod
    def from_config(cls, cfg):
        return _from_config(cfg, pretrained=False)

    def _head_outputs(self, batched_inputs):
        images = [(x["image"].to(self.device) - self.pixel_mean) / self.pixel_std for x in batched_inputs]
        batch, sizes = _batch_images(images, self.size_divisibility)
        features = self.backbone(batch)
        outputs, _ = self.sem_seg_head(features, batch.shape[-2], batch.shape[-1])
        return outputs, batched_inputs, tuple(batch.shape[-2:]), sizes


@META_ARCH_REGISTRY.register()
class PretrainedMeanShiftMaskFormer(_MetaArchBase):
    """Reference pretrained_meanshiftformer_model.py:47-497: a pretrained embedding backbone (UCN) whose unit-norm
    64-d pixel embeddings are the head's only feature map (:297-301), or any other backbone
    (``use_other_backbone``, :283-285). Images are NOT mean / std normalised here (:279-281), as in the reference."""

    @configurable
    def __init__(self, *, backbone: nn.Module, sem_seg_head: nn.Module, criterion: nn.Module, num_queries: int,
                 object_mask_threshold: float, overlap_threshold: float, metadata, size_divisibility: int,
                 sem_seg_postprocess_before_inference: bool, pixel_mean: Tuple[float], pixel_std: Tuple[float],
                 semantic_on: bool, panoptic_on: bool, instance_on: bool, test_topk_per_image: int,
                 use_embedding_loss: bool = False, embedding_loss_weight: float = 1.0, alpha: float = 0.02,
                 delta: float = 0.5, lambda_intra: float = 1.0, lambda_inter: float = 1.0, metric: str = "cosine",
                 normalize: bool = True, feature_crop: bool = False, use_depth: bool = False,
                 use_other_backbone: bool = False):
        super().__init__()
        if backbone is None:
            raise ValueError("pass the embedding backbone as `backbone`: the reference builds its UCN network "
                             "(lib/fcn/get_network_crop.py) here, which is outside this package's scope")
        self.use_other_backbone = use_other_backbone
        self.pretrained_backbone = backbone
        self.feature_crop = feature_crop
        self.use_depth = use_depth
        self._common(sem_seg_head, criterion, num_queries, object_mask_threshold, overlap_threshold, metadata,
                     size_divisibility, sem_seg_postprocess_before_inference, pixel_mean, pixel_std, semantic_on,
                     panoptic_on, instance_on, test_topk_per_image, use_embedding_loss)

    @classmethod
    def from_config(cls, cfg):
        return _from_config(cfg, pretrained=True)

    def _gather(self, batched_inputs, key):
        first = batched_inputs[0][key]
        if first.dim() == 4:                                   # one pre-batched tensor (:270-275): padded like a list
            return _batch_images(first.to(self.device), self.size_divisibility)
        return _batch_images([x[key].to(self.device) for x in batched_inputs], self.size_divisibility)

    def _head_outputs(self, batched_inputs):
        batch, sizes = self._gather(batched_inputs, "image")
        if self.use_other_backbone:
            features = self.pretrained_backbone(batch)
        else:
            if self.use_depth:
                depth, _ = self._gather(batched_inputs, "depth")
                emb = self.pretrained_backbone(batch, None, depth)
            else:
                emb = self.pretrained_backbone(batch, None)
            features = {"res5": F.normalize(emb, p=2, dim=1)}
        outputs, _ = self.sem_seg_head(features, batch.shape[-2], batch.shape[-1])
        inputs = batched_inputs if len(batched_inputs) == len(sizes) else [{} for _ in sizes]
        return outputs, inputs, tuple(batch.shape[-2:]), sizes


def _from_config(cfg, pretrained):
    """The reference's from_config (:160-251 / :140-214): needs detectron2's builders."""
    try:  # pragma: no cover - detectron2 is not part of the build image
        from detectron2.data import MetadataCatalog
        from detectron2.modeling import build_backbone, build_sem_seg_head
    except ImportError as e:
        raise NotImplementedError("from_config needs detectron2 (build_backbone / build_sem_seg_head); construct the "
                                  "model with explicit keyword arguments instead") from e
    backbone = build_backbone(cfg)
    head = build_sem_seg_head(cfg, backbone.output_shape())
    mf = cfg.MODEL.MASK_FORMER
    kw = {
        "backbone": backbone, "sem_seg_head": head,
        "criterion": build_criterion(
            head.num_classes, class_weight=mf.CLASS_WEIGHT, mask_weight=mf.MASK_WEIGHT, dice_weight=mf.DICE_WEIGHT,
            no_object_weight=mf.NO_OBJECT_WEIGHT, deep_supervision=mf.DEEP_SUPERVISION, dec_layers=mf.DEC_LAYERS,
            train_num_points=mf.TRAIN_NUM_POINTS, oversample_ratio=mf.OVERSAMPLE_RATIO,
            importance_sample_ratio=mf.IMPORTANCE_SAMPLE_RATIO),
        "num_queries": mf.NUM_OBJECT_QUERIES, "object_mask_threshold": mf.TEST.OBJECT_MASK_THRESHOLD,
        "overlap_threshold": mf.TEST.OVERLAP_THRESHOLD, "metadata": MetadataCatalog.get(cfg.DATASETS.TRAIN[0]),
        "size_divisibility": mf.SIZE_DIVISIBILITY,
        "sem_seg_postprocess_before_inference": (mf.TEST.SEM_SEG_POSTPROCESSING_BEFORE_INFERENCE or mf.TEST.PANOPTIC_ON
                                                 or mf.TEST.INSTANCE_ON),
        "pixel_mean": cfg.MODEL.PIXEL_MEAN, "pixel_std": cfg.MODEL.PIXEL_STD, "semantic_on": mf.TEST.SEMANTIC_ON,
        "instance_on": mf.TEST.INSTANCE_ON, "panoptic_on": mf.TEST.PANOPTIC_ON,
        "test_topk_per_image": cfg.TEST.DETECTIONS_PER_IMAGE,
    }
    emb = getattr(cfg.MODEL, "EMBEDDING", None)
    if emb is not None:  # reference from_config :227-236: an unsupported USE_LOSS config must fail loudly, not train without it
        kw.update(use_embedding_loss=emb.USE_LOSS, embedding_loss_weight=emb.WEIGHT_LOSS, alpha=emb.ALPHA,
                  delta=emb.DELTA, lambda_intra=emb.LAMBDA_INTRA, lambda_inter=emb.LAMBDA_INTER, metric=emb.METRIC,
                  normalize=emb.NORMALIZE)
    if pretrained:
        kw.update(feature_crop=cfg.MODEL.EMBEDDING.FEATURE_CROP, use_depth=cfg.MODEL.USE_DEPTH,
                  use_other_backbone=cfg.MODEL.USE_OTHER_BACKBONE)
    return kw

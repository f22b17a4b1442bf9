"""The BASELINE.json configurations as constructors + synthetic inputs.

``r50``   config #2: MSDeformAttnPixelDecoder (res2..res5 of a ResNet-50 at 640x480) +
          MeanShiftTransformerDecoder, 100 queries, 9 layers  (configs/mixture_ResNet50.yaml).
``ucn``   config #1/#3 stage 1: SimpleBasePixelDecoder on a unit-norm 64-d embedding map at full
          resolution + PretrainedMeanShiftTransformerDecoder, 6 layers (configs/mixture_UCN.yaml).
``crop``  config #3 stage 2: same at 224x224, 8 layers (configs/crop_mixture_UCN.yaml).
``twostage`` config #3 end to end: stage 1 (``ucn`` head) on every frame, K objects per frame cropped to 224x224,
          stage 2 (``crop`` head) on all crops of a frame in one batch, overlap test and paste-back
          (lib/fcn/test_utils.py:245-420 via fcn.test_utils / fcn.test_dataset).
``train`` config #5: the r50 head in training mode with the reference's criterion (deep supervision, 12544 points),
          synthetic rectangular ground-truth instances; one step = forward + losses + backward + clipped AdamW update.
Weights are random-init by the modules' own (reference-identical) initialisers under a fixed seed.
"""
import torch
import torch.nn.functional as F

from .d2compat import ShapeSpec

R50_SHAPES = {"res2": ShapeSpec(channels=256, stride=4), "res3": ShapeSpec(channels=512, stride=8),
              "res4": ShapeSpec(channels=1024, stride=16), "res5": ShapeSpec(channels=2048, stride=32)}

HEAD_CFG = {
    "r50": dict(pixel_decoder="MSDeformAttnPixelDecoder", decoder="MeanShiftTransformerDecoder", dec_layers=9,
                height=480, width=640),
    "ucn": dict(pixel_decoder="SimpleBasePixelDecoder", decoder="PretrainedMeanShiftTransformerDecoder",
                dec_layers=6, height=480, width=640),
    "crop": dict(pixel_decoder="SimpleBasePixelDecoder", decoder="PretrainedMeanShiftTransformerDecoder",
                 dec_layers=8, height=224, width=224),
}


def decoder_kwargs(dec_layers, num_queries=100):
    return dict(num_classes=2, hidden_dim=256, num_queries=num_queries, nheads=8, dim_feedforward=2048,
                dec_layers=dec_layers, pre_norm=False, mask_dim=256, enforce_input_project=False,
                use_meanshift_cross_attention=True, disable_attention_mask=False,
                use_meanshift_self_attention=True, decoder_block_norm=True)


def build_head(kind, seed=0):
    """-> PretrainedMeanShiftMaskFormerHead (CPU, eval) for workload ``kind``."""
    from .meanshiftformer import modeling as M
    cfg = HEAD_CFG[kind]
    torch.manual_seed(seed)
    if cfg["pixel_decoder"] == "MSDeformAttnPixelDecoder":
        shapes = R50_SHAPES
        pixel = M.MSDeformAttnPixelDecoder(shapes, transformer_dropout=0.0, transformer_nheads=8,
                                           transformer_dim_feedforward=1024, transformer_enc_layers=6, conv_dim=64,
                                           mask_dim=256, norm="GN", transformer_in_features=["res3", "res4", "res5"],
                                           common_stride=4)
    else:
        shapes = {"res5": ShapeSpec(channels=64, stride=1)}
        pixel = M.SimpleBasePixelDecoder(shapes, conv_dim=64, mask_dim=256, norm="GN")
    decoder = getattr(M, cfg["decoder"])(64, True, **decoder_kwargs(cfg["dec_layers"]))
    head = M.PretrainedMeanShiftMaskFormerHead(shapes, num_classes=2, pixel_decoder=pixel, loss_weight=1.0,
                                               ignore_value=255, transformer_predictor=decoder,
                                               transformer_in_feature="multi_scale_pixel_decoder")
    return head.eval()


def synthetic_features(kind, batch, seed=0, pin=False):
    """Backbone outputs for ``batch`` images as CPU tensors (optionally pinned)."""
    cfg = HEAD_CFG[kind]
    g = torch.Generator().manual_seed(1000 + seed)
    H, W = cfg["height"], cfg["width"]
    if kind == "r50":
        feats = {k: torch.randn(batch, s.channels, H // s.stride, W // s.stride, generator=g)
                 for k, s in R50_SHAPES.items()}
    else:  # unit-norm 64-d pixel embeddings (pretrained_meanshiftformer_model.py:298)
        feats = {"res5": F.normalize(torch.randn(batch, 64, H, W, generator=g), p=2, dim=1)}
    if pin:
        feats = {k: v.pin_memory() for k, v in feats.items()}
    return feats


MODEL_CFG = {
    # BASELINE.json config #2: R50 backbone (RGB, USE_OTHER_BACKBONE) + r50 head, 640x480
    "r50": dict(head="r50", backbone="ResNet50Features", use_depth=False, use_other_backbone=True),
    # config #1 / #3 stage 1: SEGNET RGB-D embedding (two ResNet34-8s streams, fusion add) + ucn head, 640x480
    "demo": dict(head="ucn", backbone="SegnetEmbedding", use_depth=True, use_other_backbone=False),
}


def build_model(kind, seed=0):
    """-> PretrainedMeanShiftMaskFormer (CPU, eval): backbone + head + eval tail behind the reference's META_ARCH API
    (pretrained_meanshiftformer_model.py:244-378), random-init by the modules' own initialisers under ``seed``."""
    from . import backbones
    from .meanshiftformer import PretrainedMeanShiftMaskFormer
    cfg = MODEL_CFG[kind]
    backbone = getattr(backbones, cfg["backbone"])(seed=seed)
    return PretrainedMeanShiftMaskFormer(
        backbone=backbone, sem_seg_head=build_head(cfg["head"], seed), criterion=None, num_queries=100,
        object_mask_threshold=0.8, overlap_threshold=0.8, metadata=None, size_divisibility=32,
        sem_seg_postprocess_before_inference=True, pixel_mean=[0.0, 0.0, 0.0], pixel_std=[1.0, 1.0, 1.0],
        semantic_on=False, panoptic_on=False, instance_on=True, test_topk_per_image=20, use_depth=cfg["use_depth"],
        use_other_backbone=cfg["use_other_backbone"]).eval()


def synthetic_images(kind, batch, seed=0, pin=False):
    """Model inputs for ``batch`` frames as CPU tensors: {"image": [B,3,H,W]} (already mean / std normalised - the
    Pretrained* META_ARCH does not normalise) and for the RGB-D model {"depth": [B,3,H,W]} XYZ metres
    with 5 % invalid (zero) pixels."""
    cfg = HEAD_CFG[MODEL_CFG[kind]["head"]]
    g = torch.Generator().manual_seed(8000 + seed)
    H, W = cfg["height"], cfg["width"]
    if MODEL_CFG[kind]["use_depth"]:
        out = {"image": torch.rand(batch, 3, H, W, generator=g) - 0.5}
        z = torch.rand(batch, 1, H, W, generator=g) * 1.2 + 0.3
        xy = (torch.rand(batch, 2, H, W, generator=g) - 0.5) * z
        out["depth"] = torch.cat([xy, z], 1) * (torch.rand(batch, 1, H, W, generator=g) > 0.05)
    else:
        out = {"image": torch.randn(batch, 3, H, W, generator=g)}
    if pin:
        out = {k: v.pin_memory() for k, v in out.items()}
    return out


def backbone_flops_per_image(kind):
    """Algorithmic FLOP (2*MAC) of the backbone per 640x480 frame, for reporting: torchvision ResNet-50 4.09 GMAC at
    224x224; ResNet-34 at output stride 8 (layer3 / layer4 keep the 1/8 grid) 17.33 GMAC."""
    px = 480 * 640 / (224.0 * 224.0)
    if kind == "r50":
        return 2.0 * 4.09e9 * px
    return 2.0 * 2 * 17.33e9 * px   # two ResNet34-8s streams (conv MACs counted with forward hooks: 17.33 GMAC at 224x224)


def oracle_kwargs(kind):
    cfg = HEAD_CFG[kind]
    kw = dict(pixel_decoder=cfg["pixel_decoder"], num_heads=8, dec_layers=cfg["dec_layers"])
    if cfg["pixel_decoder"] == "MSDeformAttnPixelDecoder":
        kw.update(pd_heads=8, pd_enc_layers=6)
    return kw


def head_flops_per_image(kind):
    """Algorithmic FLOP of one head forward (SURVEY.md §8d, 2*MAC), for reporting only."""
    cfg = HEAD_CFG[kind]
    Q, C, h, ffn = 100, 256, 8, 2048
    n = cfg["dec_layers"]
    if kind == "r50":
        keys = [300, 1200, 4800]
        mh = mw_ = None
        hw = 120 * 160
        S = [keys[i % 3] for i in range(n)]
        pix = 12.15e9 + 3.22e9
    else:
        hw = cfg["height"] * cfg["width"]
        S = [hw] * n
        pix = 2.0 * 9 * 64 * 256 * hw
    f = pix
    for s in S:
        f += 4.0 * Q * s * C              # QK^T + AV
        f += 2.0 * 2 * s * C * C          # K, V projections
        f += 2.0 * Q * C * C * 2          # Q, out projections
        f += 4.0 * Q * Q * C + 2.0 * Q * C * C * 4  # self attention
        f += 2.0 * 2 * Q * C * ffn        # FFN
    f += (n + 1) * (2.0 * Q * C * hw + 2.0 * Q * C * C * 3)  # mask head
    f += sum(2.0 * s * 64 * C for s in set(S))  # input_proj
    return f


class HeadTrainer(torch.nn.Module):
    """Head + criterion with the META_ARCH training contract on backbone FEATURES (the backbone is outside the hot
    path): forward({'features': dict, 'targets': [{'labels','masks'}]}) -> weighted loss dict
    (pretrained_meanshiftformer_model.py:303-334)."""

    def __init__(self, head, criterion, height, width, amp_dtype=None):
        super().__init__()
        self.sem_seg_head = head
        self.criterion = criterion
        self.size = (height, width)
        self.amp_dtype = amp_dtype  # torch.bfloat16 / torch.float16: autocast around the head, as the reference trains

    def forward(self, batch):
        dev = next(iter(batch["features"].values())).device.type
        with torch.autocast(device_type=dev, dtype=self.amp_dtype or torch.bfloat16, enabled=self.amp_dtype is not None):
            outputs, _ = self.sem_seg_head(batch["features"], *self.size)
        losses = self.criterion(outputs, batch["targets"])
        w = self.criterion.weight_dict
        return {k: v * w[k] for k, v in losses.items() if k in w}


def build_trainer(kind="r50", seed=0, amp_dtype=None):
    from .meanshiftformer.meanshiftformer_model import build_criterion
    cfg = HEAD_CFG[kind]
    head = build_head(kind, seed).train()
    crit = build_criterion(2, dec_layers=cfg["dec_layers"] + 1)  # DEC_LAYERS counts the learnable-query prediction
    return HeadTrainer(head, crit, cfg["height"], cfg["width"], amp_dtype)


def synthetic_targets(kind, batch, instances=5, seed=0):
    """Per image ``instances`` random axis-aligned rectangles (bool masks at image size) with classes in {0, 1}."""
    cfg = HEAD_CFG[kind]
    H, W = cfg["height"], cfg["width"]
    g = torch.Generator().manual_seed(2000 + seed)
    out = []
    for _ in range(batch):
        masks = torch.zeros(instances, H, W, dtype=torch.bool)
        for t in range(instances):
            h, w = int(torch.randint(H // 8, H // 3, (1,), generator=g)), int(torch.randint(W // 8, W // 3, (1,), generator=g))
            y, x = int(torch.randint(0, H - h, (1,), generator=g)), int(torch.randint(0, W - w, (1,), generator=g))
            masks[t, y:y + h, x:x + w] = True
        out.append({"labels": torch.randint(0, 2, (instances,), generator=g), "masks": masks})
    return out


class SyntheticEmbedding(torch.nn.Module):
    """Stand-in for the UCN embedding backbone (lib/networks/SEG.py, outside the hot path): a 1x1 projection of
    image (+ depth) to 64 channels, so that the two-stage pipeline runs end to end on synthetic frames."""

    def __init__(self, seed=0):
        super().__init__()
        g = torch.Generator().manual_seed(3000 + seed)
        self.weight = torch.nn.Parameter(torch.randn(64, 3, 1, 1, generator=g))

    def forward(self, image, label=None, depth=None):
        return F.conv2d(image if depth is None else image + depth, self.weight)


def build_two_stage_models(seed=0):
    """-> (stage-1 model, crop model): PretrainedMeanShiftMaskFormer wrappers around the ``ucn`` / ``crop`` heads with
    the synthetic embedding backbone, eval mode."""
    from .meanshiftformer import PretrainedMeanShiftMaskFormer
    models = []
    for i, kind in enumerate(("ucn", "crop")):
        models.append(PretrainedMeanShiftMaskFormer(
            backbone=SyntheticEmbedding(seed + i), sem_seg_head=build_head(kind, seed + i), criterion=None,
            num_queries=100, object_mask_threshold=0.8, overlap_threshold=0.8, metadata=None, size_divisibility=0,
            sem_seg_postprocess_before_inference=True, pixel_mean=[0.0, 0.0, 0.0], pixel_std=[1.0, 1.0, 1.0],
            semantic_on=False, panoptic_on=False, instance_on=True, test_topk_per_image=20, use_depth=True).eval())
    return models


def synthetic_frames(batch, objects=5, seed=0, height=480, width=640):
    """RGB-D frames + a stage-1 label map with ``objects`` disjoint rectangles per frame (ids 2..objects+1, the ids
    get_confident_instances produces, lib/fcn/test_utils.py:35-52). Random-init networks segment noise, so the bench
    substitutes this map for stage 1's output to fix the number of crops (K = 5)."""
    g = torch.Generator().manual_seed(4000 + seed)
    image = torch.rand(batch, 3, height, width, generator=g) - 0.5
    depth = torch.rand(batch, 3, height, width, generator=g) * 1.2 + 0.3
    labels = torch.zeros(batch, height, width)
    cell = width // objects
    for b in range(batch):
        for k in range(objects):
            h = int(torch.randint(height // 6, height // 2, (1,), generator=g))
            w = int(torch.randint(cell // 3, cell - 8, (1,), generator=g))
            y = int(torch.randint(4, height - h - 4, (1,), generator=g))
            x = k * cell + int(torch.randint(2, cell - w - 2, (1,), generator=g))
            labels[b, y:y + h, x:x + w] = k + 2
    return image, depth, labels

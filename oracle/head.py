"""Oracle: segmentation head glue and the eval-mode tail. Test infrastructure only.

Follows MSMFormer/meanshiftformer/modeling/meta_arch/meanshift_former_head.py:246-275 and
MSMFormer/meanshiftformer/pretrained_meanshiftformer_model.py:335-343, 461-497.
"""
import torch
import torch.nn.functional as F

from .decoder import decoder_forward
from .pixel_decoder import msdeform_pixel_decoder_forward, simple_pixel_decoder_forward


def _sub(sd, prefix):
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}


def head_forward(sd, features, *, pixel_decoder, num_heads, dec_layers, pd_heads=None, pd_enc_layers=None):
    """PretrainedMeanShiftMaskFormerHead.layers (meanshift_former_head.py:246-275):
    pixel decoder -> (mask_features, _, multi_scale_features) -> transformer decoder.
    ``sd`` uses the head's key names (``pixel_decoder.*``, ``predictor.*``).
    Returns (predictions, mask_features)."""
    psd = _sub(sd, "pixel_decoder.")
    if pixel_decoder == "MSDeformAttnPixelDecoder":
        mf, _, ms = msdeform_pixel_decoder_forward(psd, features, n_heads=pd_heads, enc_layers=pd_enc_layers)
    elif pixel_decoder == "SimpleBasePixelDecoder":
        mf, _, ms = simple_pixel_decoder_forward(psd, features)
    else:
        raise ValueError(pixel_decoder)
    out = decoder_forward(_sub(sd, "predictor."), ms, mf, num_heads=num_heads, num_layers=dec_layers)
    return out, mf


def instance_inference(mask_cls, mask_pred, num_classes, topk):
    """pretrained_meanshiftformer_model.py:461-497 for one image (panoptic_on False).

    mask_cls [Q,K+1], mask_pred [Q,H,W] logits already at image size. Returns dict with
    pred_masks float {0,1} [topk,H,W], pred_boxes [topk,4] (x0,y0,x1,y1; BitMasks.get_bounding_boxes:
    tight box with exclusive max, zeros for empty masks), scores, pred_classes, query_index."""
    Q = mask_cls.shape[0]
    scores = F.softmax(mask_cls, dim=-1)[:, :-1]
    labels = torch.arange(num_classes).unsqueeze(0).repeat(Q, 1).flatten(0, 1)
    s, idx = scores.flatten(0, 1).topk(topk, sorted=False)
    cls = labels[idx]
    qidx = idx // num_classes
    m = mask_pred[qidx]
    binm = (m > 0).float()
    boxes = torch.zeros(topk, 4)
    for i in range(topk):
        ys, xs = torch.where(m[i] > 0)
        if len(xs) > 0:
            boxes[i] = torch.tensor([xs.min(), ys.min(), xs.max() + 1, ys.max() + 1], dtype=torch.float32)
    mscore = (m.sigmoid().flatten(1) * binm.flatten(1)).sum(1) / (binm.flatten(1).sum(1) + 1e-6)
    return {"pred_masks": binm, "pred_boxes": boxes, "scores": s * mscore, "pred_classes": cls, "query_index": qidx}


def eval_tail(outputs, image_size, num_classes, topk):
    """pretrained_meanshiftformer_model.py:335-378 with output size == padded image size
    (sem_seg_postprocess is then the identity): upsample all mask logits, per-image instance_inference."""
    up = F.interpolate(outputs["pred_masks"], size=image_size, mode="bilinear", align_corners=False)
    return [instance_inference(c, m, num_classes, topk) for c, m in zip(outputs["pred_logits"], up)]

"""Oracle: eval-mode tail (mask upsample + instance_inference). Test infrastructure only.

Follows MSMFormer/meanshiftformer/pretrained_meanshiftformer_model.py:337-343 (bilinear upsample of every
predicted mask) and :461-497 (instance_inference, panoptic_on=False); meanshiftformer_model.py:289-295, 414-450 are
the same lines. ``BitMasks.get_bounding_boxes`` is detectron2 (third-party, not vendored in the reference;
detectron2 v0.6 structures/masks.py): per mask, the first/last column and row holding a True, as
(x0, y0, x1 + 1, y1 + 1), zeros for an empty mask.
"""
import torch
import torch.nn.functional as F


def get_bounding_boxes(bitmasks):
    boxes = torch.zeros(bitmasks.shape[0], 4, dtype=torch.float32)
    x_any = torch.any(bitmasks, dim=1)
    y_any = torch.any(bitmasks, dim=2)
    for i in range(bitmasks.shape[0]):
        x = torch.where(x_any[i, :])[0]
        y = torch.where(y_any[i, :])[0]
        if len(x) > 0 and len(y) > 0:
            boxes[i, :] = torch.as_tensor([x[0], y[0], x[-1] + 1, y[-1] + 1], dtype=torch.float32)
    return boxes


def instance_inference(mask_cls, mask_pred, num_classes, topk):
    """:461-497 on one image: mask_cls [Q,K+1], mask_pred [Q,H,W] (already at full resolution).
    Returns dict(pred_masks, pred_boxes, scores, pred_classes, query_index) in topk's order."""
    Q = mask_cls.shape[0]
    scores = F.softmax(mask_cls, dim=-1)[:, :-1]
    labels = torch.arange(num_classes).unsqueeze(0).repeat(Q, 1).flatten(0, 1)
    scores_per_image, topk_indices = scores.flatten(0, 1).topk(topk, sorted=False)
    labels_per_image = labels[topk_indices]
    topk_indices = topk_indices // num_classes
    mask_pred = mask_pred[topk_indices]
    pred_masks = (mask_pred > 0).float()
    boxes = get_bounding_boxes(mask_pred > 0)
    mask_scores = (mask_pred.sigmoid().flatten(1) * pred_masks.flatten(1)).sum(1) / (pred_masks.flatten(1).sum(1) + 1e-6)
    return {"pred_masks": pred_masks, "pred_boxes": boxes, "scores": scores_per_image * mask_scores,
            "pred_classes": labels_per_image, "query_index": topk_indices}


def inference_tail(pred_logits, pred_masks, image_size, topk):
    """:335-378 with instance_on only and output size == input size (sem_seg_postprocess = identity)."""
    up = F.interpolate(pred_masks, size=tuple(image_size), mode="bilinear", align_corners=False)
    K = pred_logits.shape[-1] - 1
    return [instance_inference(c, m, K, topk) for c, m in zip(pred_logits, up)]


def canonical(fields, num_classes):
    """rows in a canonical order (ascending flattened (query, class) index) - topk(sorted=False) has none."""
    order = torch.argsort(fields["query_index"] * num_classes + fields["pred_classes"])
    return {k: v[order] for k, v in fields.items()}


def get_confident_instances(fields, topk=False, score=0.7, num_class=2, low_threshold=0.4):
    """lib/fcn/test_utils.py:35-52 on a dict of per-instance tensors."""
    if topk:
        if num_class < 2:
            return fields
        keep = (fields["pred_classes"] == 1) & (fields["scores"] > low_threshold)
    else:
        keep = fields["scores"] > score
    return {k: v[keep] for k, v in fields.items()}


def combine_masks(fields):
    """lib/fcn/test_utils.py:93-112: instance i paints label i + 2 over everything painted before."""
    masks = fields["pred_masks"]
    out = torch.zeros(masks.shape[-2:], dtype=torch.float64)
    for i in range(masks.shape[0]):
        out[masks[i] != 0] = i + 2
    return out

"""Instances -> label map, the first half of the reference's test-time glue (lib/fcn/test_utils.py:35-52, 93-112).

``get_confident_instances`` / ``combine_masks`` keep the reference's names and semantics on the tensor form of
``Instances`` that ``meanshiftformer.instance_inference`` returns (a dict with pred_masks / pred_boxes / scores /
pred_classes). ``label_map_from_outputs`` is the fused form the pipeline actually needs: decoder outputs -> label map,
without the [T, H, W] float masks ever being written (the reference writes them, copies them to the host and loops
over them in numpy): top-k, mask scores from one reduction pass, then one pass that labels every pixel.
"""
import torch

from .. import ops


def get_confident_instances(outputs, topk=False, score=0.7, num_class=2, low_threshold=0.4):
    """:35-52 on one image's fields (dict of tensors, rows = instances)."""
    inst = outputs["instances"] if "instances" in outputs else outputs
    if topk:
        if num_class < 2:
            return inst
        keep = (inst["pred_classes"] == 1) & (inst["scores"] > low_threshold)
    else:
        keep = inst["scores"] > score
    return {k: v[keep] for k, v in inst.items()}


def combine_masks(instances):
    """:93-112 - [N,H,W] binary masks -> [H,W] label map, instance i gets label i + 2, later instances overwrite
    earlier ones. Stays on the device (the reference returns a numpy float64 array)."""
    masks = instances["pred_masks"]
    num, H, W = masks.shape
    if num == 0:
        return torch.zeros(H, W, device=masks.device)
    ids = torch.arange(2, 2 + num, device=masks.device, dtype=torch.float32).view(num, 1, 1)
    return (ids * (masks != 0)).max(dim=0)[0]   # labels increase with the instance index: last writer = largest label


def label_map_from_outputs(pred_logits, pred_masks, image_size, test_topk_per_image, topk=False, score=0.7,
                           num_class=2, low_threshold=0.4):
    """instance_inference -> get_confident_instances -> combine_masks for a batch, fused:
    pred_logits [B,Q,K+1], pred_masks [B,Q,h,w] -> (label_map fp32 [B,H,W], dict of the per-instance fields incl.
    ``instance_label`` int32 [B,T]: the label each kept instance received, -1 = dropped)."""
    query, cls, cls_score = ops.instance_topk(pred_logits, test_topk_per_image)
    _, boxes, scores = ops.instance_masks(pred_masks, query, cls_score, image_size, want_masks=False)
    label_map, inst_label = ops.instance_label_map(pred_masks, query, scores, cls, image_size, topk_mode=topk,
                                                   num_class=num_class, score=score, low_threshold=low_threshold)
    return label_map, {"pred_boxes": boxes, "scores": scores, "pred_classes": cls, "query_index": query,
                       "instance_label": inst_label}


def two_stage_label_maps(model, model_crop, image, depth=None, *, topk=False, confident_score=0.7,
                         low_threshold=0.4, depth_threshold=0.5, crop_size=224):
    """The inference part of ``test_sample_crop`` (lib/fcn/test_utils.py:245-420) for one frame, on the device:
    stage 1 label map -> depth filter -> padded ROI crops -> stage 2 on ALL crops in one batched forward (the
    reference calls its crop predictor once per object, :397-405) -> overlap test, ordering and paste-back.
    ``model`` / ``model_crop``: META_ARCH wrappers of this package (``label_maps``); image [1,3,H,W] (or [3,H,W]),
    depth [1,3,H,W] or None. Returns (out_label [1,H,W], out_label_refined [1,H,W] or None)."""
    from . import test_dataset as td
    if image.dim() == 3:
        image = image.unsqueeze(0)
    if depth is not None and depth.dim() == 3:
        depth = depth.unsqueeze(0)
    sample = {"image": image[0]} if depth is None else {"image": image[0], "depth": depth[0]}
    out_label, _ = model.label_maps([sample], topk=topk, score=confident_score, low_threshold=low_threshold)
    if depth is not None:
        out_label = td.filter_labels_depth(out_label, depth, depth_threshold)
    refined = None
    if model_crop is not None:
        rgb_crop, out_label_crop, rois, depth_crop = td.crop_rois(image, out_label.clone(), depth, crop_size=crop_size)
        if rgb_crop.shape[0] > 0:
            crops = {"image": rgb_crop} if depth_crop is None else {"image": rgb_crop, "depth": depth_crop}
            labels_crop, _ = model_crop.label_maps([crops], topk=topk, score=confident_score,
                                                   low_threshold=low_threshold)
            refined, _ = td.match_label_crop(out_label, labels_crop, out_label_crop, rois, depth_crop)
    return out_label, refined

"""Eval-mode tail of the reference's meta-architectures, on the device and for the kept queries only.

Reference: ``PretrainedMeanShiftMaskFormer.forward`` (eval branch) and ``instance_inference``,
MSMFormer/meanshiftformer/pretrained_meanshiftformer_model.py:335-378, 461-497 (identical code in
meanshiftformer_model.py:287-330, 414-450). The reference upsamples all ``num_queries`` masks of every image to
the input resolution and then keeps ``test_topk_per_image`` of them; here the top-k comes first and one CUDA pass
writes the binary masks, boxes and scores of the kept queries (``ops.instance_topk`` + ``ops.instance_masks``).

Same field names as the reference's ``Instances``: pred_masks, pred_boxes, scores, pred_classes. With detectron2
installed ``to_instances`` wraps them into real ``Instances`` / ``Boxes`` objects; without it they stay tensors.
"""
from .. import ops


def instance_inference_batched(pred_logits, pred_masks, image_size, topk):
    """pred_logits [B,Q,K+1], pred_masks [B,Q,h,w] (decoder outputs, low resolution), image_size (H, W) ->
    dict(pred_masks [B,T,H,W] 0/1 float, pred_boxes [B,T,4] XYXY, scores [B,T], pred_classes int64 [B,T],
    query_index int64 [B,T]). Rows are ordered by descending class score (the reference's
    ``topk(sorted=False)`` promises no order). panoptic_on=False, as in every UOIS config."""
    query, cls, cls_score = ops.instance_topk(pred_logits, topk)
    masks, boxes, scores = ops.instance_masks(pred_masks, query, cls_score, image_size)
    return {"pred_masks": masks, "pred_boxes": boxes, "scores": scores, "pred_classes": cls, "query_index": query}


def instance_inference(mask_cls, mask_pred, image_size, topk):
    """One image, the reference's argument order: mask_cls [Q,K+1], mask_pred [Q,h,w] low-resolution logits
    (the reference receives them already upsampled; upsampling is part of this call)."""
    r = instance_inference_batched(mask_cls.unsqueeze(0), mask_pred.unsqueeze(0), image_size, topk)
    return {k: v[0] for k, v in r.items()}


def inference_tail(outputs, image_size, topk):
    """The eval branch of ``forward`` with instance_on only (sem_seg_postprocess is the identity when the output
    size equals the padded input size, which is how the UOIS test scripts call it, lib/fcn/test_utils.py:93-112):
    returns the reference's list of {"instances": ...}, one entry per image."""
    r = instance_inference_batched(outputs["pred_logits"], outputs["pred_masks"], image_size, topk)
    B = outputs["pred_logits"].shape[0]
    return [{"instances": to_instances({k: v[b] for k, v in r.items()}, image_size)} for b in range(B)]


def to_instances(fields, image_size):
    try:  # pragma: no cover - detectron2 is not part of the build image
        from detectron2.structures import Boxes, Instances
    except ImportError:
        return fields
    inst = Instances(tuple(image_size))
    inst.pred_masks = fields["pred_masks"]
    inst.pred_boxes = Boxes(fields["pred_boxes"])
    inst.scores = fields["scores"]
    inst.pred_classes = fields["pred_classes"]
    return inst

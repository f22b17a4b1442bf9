"""Training step around the META_ARCH wrappers (BASELINE.json config #5).

The reference trains through detectron2's DefaultTrainer (tabletop_train_net_pretrained.py:270-330): AdamW, full-model
gradient clipping (SOLVER.CLIP_GRADIENTS, clip value 0.01, configs/mixture_ResNet50.yaml), DDP across GPUs with one
process per GPU. That loop is outside the hot path; what is kept here is the step itself, so that the bench and
the tests can drive forward + losses + backward + update the way the reference's trainer does:

* one process per GPU, ``torch.nn.parallel.DistributedDataParallel`` over NCCL (NVLink / NVSwitch): the gradient
  all-reduce is the ONLY collective of the data path besides the criterion's one-float ``num_masks`` all-reduce
  (criterion.py:225-227); buckets are sized for launch latency and overlap with the backward, not for link count;
* losses come back as device scalars - nothing in the step reads the device, apart from the matcher's single copy
  of the cost matrices (modeling/matcher.py).
"""
import torch
import torch.distributed as dist
from torch import nn

from . import ops


def wrap_ddp(model, local_rank=None, bucket_cap_mb=64):
    """DistributedDataParallel when torch.distributed is initialised with more than one rank, else the model itself.
    ``local_rank`` = CUDA device of this process (None on CPU / gloo)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return model
    ids = None if local_rank is None else [local_rank]
    return nn.parallel.DistributedDataParallel(model, device_ids=ids, bucket_cap_mb=bucket_cap_mb,
                                               gradient_as_bucket_view=True, broadcast_buffers=False)


def build_optimizer(model, lr=1e-4, weight_decay=0.05, backbone_multiplier=0.1, weight_decay_norm=0.0,
                    weight_decay_embed=0.0):
    """AdamW with the reference's per-parameter hyper-parameters (tabletop_train_net_pretrained.py:113-160: modules
    whose name contains "backbone" train at lr x BACKBONE_MULTIPLIER, normalisation layers and embeddings have their
    own weight decay; defaults = configs/Base-COCO-InstanceSegmentation.yaml:21-28). The reference makes one
    parameter group per tensor; here tensors with equal hyper-parameters share a group, so the update is a handful
    of multi-tensor launches (fused AdamW on CUDA) instead of one launch chain per tensor - same arithmetic."""
    norm_types = (nn.BatchNorm1d, nn.BatchNorm2d, nn.BatchNorm3d, nn.SyncBatchNorm, nn.GroupNorm, nn.InstanceNorm1d,
                  nn.InstanceNorm2d, nn.InstanceNorm3d, nn.LayerNorm, nn.LocalResponseNorm)
    groups, seen = {}, set()
    for mod_name, module in model.named_modules():
        for name, p in module.named_parameters(recurse=False):
            if not p.requires_grad or p in seen:
                continue
            seen.add(p)
            p_lr = lr * backbone_multiplier if "backbone" in mod_name else lr
            wd = weight_decay
            if "relative_position_bias_table" in name or "absolute_pos_embed" in name:
                wd = 0.0
            if isinstance(module, norm_types):
                wd = weight_decay_norm
            if isinstance(module, nn.Embedding):
                wd = weight_decay_embed
            groups.setdefault((p_lr, wd), []).append(p)
    param_groups = [{"params": ps, "lr": k[0], "weight_decay": k[1]} for k, ps in groups.items()]
    on_cuda = all(p.is_cuda for ps in groups.values() for p in ps)
    return torch.optim.AdamW(param_groups, lr=lr, fused=True if on_cuda else None)


def train_step(model, optimizer, batched_inputs, clip_value=0.01):
    """One optimisation step: weighted loss dict -> backward (DDP all-reduces the gradients bucket by bucket while
    the backward is still running) -> full-model gradient-norm clipping -> AdamW update. Returns the detached loss
    dict (device scalars; read them when you need them)."""
    losses = model(batched_inputs)
    total = sum(losses.values())
    optimizer.zero_grad(set_to_none=True)
    total.backward()
    if clip_value and clip_value > 0:
        params = [p for g in optimizer.param_groups for p in g["params"] if p.grad is not None]
        nn.utils.clip_grad_norm_(params, clip_value)
    optimizer.step()
    # fused / multi-tensor optimizers update parameters in place WITHOUT bumping tensor._version: every cache of
    # derived weights (prepared 16-bit copies, concatenated K/V weights, row-bias tables) is keyed by this epoch
    ops.bump_weights_epoch()
    return {k: v.detach() for k, v in losses.items()}

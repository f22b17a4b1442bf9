"""MaskFormer set criterion - mirror of the reference's modeling/criterion.py (``SetCriterion``: same constructor,
``forward(outputs, targets) -> {loss_ce, loss_mask, loss_dice, loss_*_{i}}``, same ``empty_weight`` buffer), organised
so that a training step synchronises with the host ONCE:

* the matching of the final and of every auxiliary prediction is one call (``HungarianMatcher.match_layers``: one
  device->host copy of all cost matrices; the reference synchronises per image and per layer, criterion.py:218,238);
* the number of ground-truth masks is known on the host from the target shapes, so it needs no ``.item()``
  (criterion.py:221-227 reads it back from the device); under torch.distributed it stays a device scalar;
* the point-sampled mask losses of all layers are reduced together.
Point sampling (importance sampling on -|logit|) follows detectron2's PointRend helpers, see point_features.py.
"""
import torch
import torch.distributed as dist
import torch.nn.functional as F
from torch import nn

from .matcher import pad_targets
from .point_features import PointSource, point_sample, uncertain_point_coords


def dice_loss(inputs, targets, num_masks):
    """criterion.py:21-41, for [..., N, P] point logits / labels: summed over the N masks, / num_masks."""
    inputs = inputs.sigmoid()
    numerator = 2 * (inputs * targets).sum(-1)
    denominator = inputs.sum(-1) + targets.sum(-1)
    return (1 - (numerator + 1) / (denominator + 1)).sum(-1) / num_masks


def sigmoid_ce_loss(inputs, targets, num_masks):
    """criterion.py:49-65, same layout as dice_loss."""
    return F.binary_cross_entropy_with_logits(inputs, targets, reduction="none").mean(-1).sum(-1) / num_masks


class SetCriterion(nn.Module):
    """Reference criterion.py:93-263."""

    def __init__(self, num_classes, matcher, weight_dict, eos_coef, losses, num_points, oversample_ratio,
                 importance_sample_ratio):
        super().__init__()
        self.num_classes = num_classes
        self.matcher = matcher
        self.weight_dict = weight_dict
        self.eos_coef = eos_coef
        self.losses = losses
        empty_weight = torch.ones(self.num_classes + 1)
        empty_weight[-1] = self.eos_coef
        self.register_buffer("empty_weight", empty_weight)
        self.num_points = num_points
        self.oversample_ratio = oversample_ratio
        self.importance_sample_ratio = importance_sample_ratio

    # ---------------------------------------------------------------------------------------------- pieces
    @staticmethod
    def _permutation_idx(indices, which):
        batch_idx = torch.cat([torch.full_like(pair[which], b) for b, pair in enumerate(indices)])
        return batch_idx, torch.cat([pair[which] for pair in indices])

    @staticmethod
    def _device_indices(layer_indices, device):
        """indices[layer][image] = (pred idx, target idx) on the host -> ONE device tensor [3, L, N]: image, query and
        target (column of the padded targets) of every matched pair. N = matched pairs per layer - the same for every
        layer, since every target of an image is matched exactly once. The reference moves one small index tensor per
        image, layer and use; the step is host-bound, so this is one copy per step."""
        rows = []
        for indices in layer_indices:
            b = torch.cat([torch.full_like(i, k) for k, (i, _) in enumerate(indices)])
            rows.append(torch.stack([b, torch.cat([i for i, _ in indices]), torch.cat([j for _, j in indices])]))
        return torch.stack(rows, 1).to(device)    # [3, L, N]

    def loss_labels(self, outputs, targets, indices, num_masks):
        """criterion.py:124-141: weighted cross entropy, unmatched queries -> the no-object class (one layer)."""
        src_logits = outputs["pred_logits"].float()
        dev = src_logits.device
        b_idx, s_idx = self._permutation_idx(indices, 0)
        matched = torch.cat([t["labels"][J.to(t["labels"].device)] for t, (_, J) in zip(targets, indices)])
        target_classes = torch.full(src_logits.shape[:2], self.num_classes, dtype=torch.int64, device=dev)
        target_classes[b_idx.to(dev), s_idx.to(dev)] = matched.to(dev)
        return {"loss_ce": F.cross_entropy(src_logits.transpose(1, 2), target_classes, self.empty_weight)}

    def _label_losses(self, layer_outputs, labels_pad, idx):
        """loss_labels for all layers at once -> [L]. F.cross_entropy's weighted mean is sum(w_t * nll) / sum(w_t) over
        the layer's B * Q queries; computed per layer from one unreduced call."""
        logits = torch.stack([o["pred_logits"].float() for o in layer_outputs])          # [L,B,Q,K+1]
        L, B, Q, K1 = logits.shape
        target = torch.full((L, B, Q), self.num_classes, dtype=torch.int64, device=logits.device)
        if idx.shape[2]:
            l_idx = torch.arange(L, device=logits.device)[:, None].expand(L, idx.shape[2])
            target[l_idx, idx[0], idx[1]] = labels_pad[idx[0], idx[2]]
        weight = self.empty_weight.to(logits.dtype)
        nll = F.cross_entropy(logits.reshape(L * B * Q, K1), target.reshape(-1), weight, reduction="none").view(L, B * Q)
        return nll.sum(1) / weight[target].view(L, B * Q).sum(1)

    def _mask_losses(self, layer_outputs, masks_pad, idx, num_masks, point_source):
        """criterion.py:143-189 for all layers: {'loss_mask': [layers], 'loss_dice': [layers]}."""
        dev = layer_outputs[0]["pred_masks"].device
        nl = len(layer_outputs)
        N = idx.shape[2]
        if N == 0:  # no ground truth in the whole batch: the mask losses vanish but stay attached to the graph
            zero = torch.stack([o["pred_masks"].sum() * 0.0 for o in layer_outputs])
            return {"loss_mask": zero, "loss_dice": zero.clone()}
        src = [out["pred_masks"][idx[0, l], idx[1, l]] for l, out in enumerate(layer_outputs)]   # L x [N,h,w]
        P = self.num_points
        num_sampled = int(P * self.oversample_ratio)
        num_random = P - int(self.importance_sample_ratio * P)
        candidates = point_source.oversampled_points(nl, N, num_sampled, dev)
        fill = point_source.random_points(nl, N, num_random, dev)
        src_all = torch.stack(src).flatten(0, 1)[:, None]  # [layers * N, 1, h, w]
        with torch.no_grad():
            coords = uncertain_point_coords(src_all.float(), candidates.flatten(0, 1), fill.flatten(0, 1), P,
                                            self.importance_sample_ratio)
            tgt_all = masks_pad[idx[0].reshape(-1), idx[2].reshape(-1)]                            # [layers * N, H, W]
            labels = point_sample(tgt_all[:, None].to(src_all.dtype), coords, align_corners=False).squeeze(1)
            labels = labels.unflatten(0, (nl, N))                                                  # [layers, N, P]
        logits = point_sample(src_all, coords, align_corners=False).squeeze(1).unflatten(0, (nl, N))
        return {"loss_mask": sigmoid_ce_loss(logits, labels, num_masks), "loss_dice": dice_loss(logits, labels, num_masks)}

    # ---------------------------------------------------------------------------------------------- forward
    def forward(self, outputs, targets, point_source=None):
        """outputs: {'pred_logits' [B,Q,K+1], 'pred_masks' [B,Q,h,w], 'aux_outputs': [same, ...]};
        targets: list of {'labels' [T_b], 'masks' [T_b,H,W]}. Keys as the reference: the final prediction's losses
        are unsuffixed, auxiliary layer i gets ``_{i}`` (criterion.py:233-245)."""
        src = point_source if point_source is not None else PointSource()
        layer_outputs = [{k: v for k, v in outputs.items() if k != "aux_outputs"}] + list(outputs.get("aux_outputs", []))
        suffix = [""] + [f"_{i}" for i in range(len(layer_outputs) - 1)]
        dev = layer_outputs[0]["pred_logits"].device
        padded = pad_targets(targets, dev)          # labels [B,Tmax], masks [B,Tmax,H,W]: shared with the matcher
        layer_indices = self.matcher.match_layers(layer_outputs, targets, src, padded=padded)
        idx = self._device_indices(layer_indices, dev)

        num_masks = float(sum(int(t["labels"].shape[0]) for t in targets))
        if dist.is_available() and dist.is_initialized():  # average over the ranks, criterion.py:225-227
            n = torch.as_tensor([num_masks], dtype=torch.float, device=dev)
            dist.all_reduce(n)
            num_masks = torch.clamp(n / dist.get_world_size(), min=1)[0]
        else:
            num_masks = max(num_masks, 1.0)

        unknown = [l for l in self.losses if l not in ("labels", "masks")]
        assert not unknown, f"do you really want to compute {unknown[0]} loss?"
        losses = {}
        mask_losses = label_losses = None
        if "masks" in self.losses:
            mask_losses = self._mask_losses(layer_outputs, padded[1], idx, num_masks, src)
        if "labels" in self.losses:
            label_losses = self._label_losses(layer_outputs, padded[0], idx)
        for l, sfx in enumerate(suffix):
            for name in self.losses:  # the reference's key order: per layer, in the order of self.losses
                if name == "labels":
                    losses["loss_ce" + sfx] = label_losses[l]
                else:
                    losses["loss_mask" + sfx] = mask_losses["loss_mask"][l]
                    losses["loss_dice" + sfx] = mask_losses["loss_dice"][l]
        return losses

    def __repr__(self):
        head = "Criterion " + self.__class__.__name__
        body = [f"matcher: {self.matcher.__repr__(_repr_indent=8)}", f"losses: {self.losses}",
                f"weight_dict: {self.weight_dict}", f"num_classes: {self.num_classes}", f"eos_coef: {self.eos_coef}",
                f"num_points: {self.num_points}", f"oversample_ratio: {self.oversample_ratio}",
                f"importance_sample_ratio: {self.importance_sample_ratio}"]
        return "\n".join([head] + [" " * 4 + line for line in body])

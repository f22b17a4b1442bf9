"""Hungarian matching between predictions and ground-truth masks - mirror of the reference's modeling/matcher.py
(``HungarianMatcher``, same constructor and ``forward(outputs, targets)`` contract), organised for the device:

The reference walks the batch in Python and moves one cost matrix to the host per image and per decoder layer
((layers + 1) x B device->host syncs per step, matcher.py:104-157). Here the cost matrices of ALL images of a layer
come from three batched contractions on padded targets, the layers of a step are stacked, and ONE copy brings
[layers, B, Q, Tmax] floats to the host, where scipy's ``linear_sum_assignment`` (the reference's solver,
matcher.py:8,151) runs on each [Q, T_b] block.
"""
import torch
import torch.nn.functional as F
from scipy.optimize import linear_sum_assignment
from torch import nn

from .point_features import PointSource, point_sample


def pad_targets(targets, device, dtype=torch.float32):
    """list of {'labels' [T_b], 'masks' [T_b,H,W]} -> (labels int64 [B,Tmax] (0 padded), masks [B,Tmax,H,W]
    (0 padded), counts list[int])."""
    counts = [int(t["labels"].shape[0]) for t in targets]
    tmax = max(max(counts), 1)
    H, W = targets[0]["masks"].shape[-2:]
    labels = torch.zeros(len(targets), tmax, dtype=torch.int64, device=device)
    masks = torch.zeros(len(targets), tmax, H, W, dtype=dtype, device=device)
    for b, t in enumerate(targets):
        if counts[b]:
            labels[b, :counts[b]] = t["labels"].to(device)
            masks[b, :counts[b]] = t["masks"].to(device=device, dtype=dtype)
    return labels, masks, counts


class HungarianMatcher(nn.Module):
    """Reference matcher.py:71-189. cost = cost_mask * sigmoid-CE + cost_class * (-prob[target class])
    + cost_dice * dice, masks compared on ``num_points`` points shared by all masks of an image."""

    def __init__(self, cost_class: float = 1, cost_mask: float = 1, cost_dice: float = 1, num_points: int = 0):
        super().__init__()
        self.cost_class = cost_class
        self.cost_mask = cost_mask
        self.cost_dice = cost_dice
        assert cost_class != 0 or cost_mask != 0 or cost_dice != 0, "all costs cant be 0"
        self.num_points = num_points

    @torch.no_grad()
    def cost_matrices(self, outputs, labels, tgt_points, coords):
        """One layer: outputs {'pred_logits' [B,Q,K+1], 'pred_masks' [B,Q,h,w]}, labels [B,Tmax], tgt_points
        [B,Tmax,P] (ground-truth masks sampled at ``coords`` [B,P,2]) -> cost [B,Q,Tmax] (padded columns are
        meaningless). matcher.py:107-149."""
        prob = outputs["pred_logits"].float().softmax(-1)
        B, Q = prob.shape[:2]
        cost_class = -torch.gather(prob, 2, labels.unsqueeze(1).expand(B, Q, labels.shape[1]))
        out_pts = point_sample(outputs["pred_masks"].float(), coords, align_corners=False)  # [B,Q,P]
        P = out_pts.shape[-1]
        tgt_t = tgt_points.transpose(1, 2)  # [B,P,Tmax]
        pos = F.binary_cross_entropy_with_logits(out_pts, torch.ones_like(out_pts), reduction="none")
        neg = F.binary_cross_entropy_with_logits(out_pts, torch.zeros_like(out_pts), reduction="none")
        cost_mask = (torch.bmm(pos, tgt_t) + torch.bmm(neg, 1 - tgt_t)) / P
        sig = out_pts.sigmoid()
        numerator = 2 * torch.bmm(sig, tgt_t)
        denominator = sig.sum(-1)[:, :, None] + tgt_points.sum(-1)[:, None, :]
        cost_dice = 1 - (numerator + 1) / (denominator + 1)
        return self.cost_mask * cost_mask + self.cost_class * cost_class + self.cost_dice * cost_dice

    @torch.no_grad()
    def cost_matrices_layers(self, layer_outputs, labels, masks, coords):
        """All decoder layers of a step at once (the training step is host-bound: ~15 launches per layer become ~25
        per step). layer_outputs: L dicts; labels [B,Tmax]; masks [B,Tmax,H,W] padded ground truth; coords [L,B,P,2]
        -> cost [L,B,Q,Tmax], the same arithmetic as ``cost_matrices`` per layer."""
        L, B, P = coords.shape[:3]
        T = labels.shape[1]
        prob = torch.stack([o["pred_logits"].float() for o in layer_outputs]).softmax(-1)      # [L,B,Q,K+1]
        Q = prob.shape[2]
        cost_class = -torch.gather(prob, 3, labels[None, :, None, :].expand(L, B, Q, T))
        # ground truth at every layer's points with ONE sampling call: the points of all layers side by side per image
        tgt_points = point_sample(masks, coords.permute(1, 0, 2, 3).reshape(B, L * P, 2), align_corners=False)
        tgt_points = tgt_points.view(B, T, L, P).permute(2, 0, 1, 3).reshape(L * B, T, P)      # [L*B,Tmax,P]
        out_pts = torch.stack([point_sample(o["pred_masks"].float(), coords[l], align_corners=False)
                               for l, o in enumerate(layer_outputs)]).flatten(0, 1)           # [L*B,Q,P]
        tgt_t = tgt_points.transpose(1, 2)
        pos = F.binary_cross_entropy_with_logits(out_pts, torch.ones_like(out_pts), reduction="none")
        neg = F.binary_cross_entropy_with_logits(out_pts, torch.zeros_like(out_pts), reduction="none")
        cost_mask = (torch.bmm(pos, tgt_t) + torch.bmm(neg, 1 - tgt_t)) / P
        sig = out_pts.sigmoid()
        numerator = 2 * torch.bmm(sig, tgt_t)
        denominator = sig.sum(-1)[:, :, None] + tgt_points.sum(-1)[:, None, :]
        cost_dice = 1 - (numerator + 1) / (denominator + 1)
        cost = self.cost_mask * cost_mask + self.cost_dice * cost_dice
        return cost.view(L, B, Q, T) + self.cost_class * cost_class

    @torch.no_grad()
    def match_layers(self, layer_outputs, targets, point_source=None, padded=None):
        """Matching for every decoder layer of a step with one device->host copy.
        layer_outputs: list of {'pred_logits','pred_masks'}; returns indices[layer][image] = (pred idx, target idx)
        int64 CPU tensors, as the reference's forward returns per layer. ``padded``: pad_targets(targets) when the
        caller already has it."""
        dev = layer_outputs[0]["pred_logits"].device
        B = len(targets)
        labels, masks, counts = padded if padded is not None else pad_targets(targets, dev)
        src = point_source if point_source is not None else PointSource()
        tf32 = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False  # near-tied costs must not flip with the GEMM mode
        try:
            coords = src.matcher_points(len(layer_outputs), B, self.num_points, dev)  # [layers,B,P,2]
            with torch.autocast(device_type=dev.type, enabled=False):  # fp32 costs, as matcher.py:134-141
                costs = self.cost_matrices_layers(layer_outputs, labels, masks, coords)
            host = costs.cpu()  # the step's only synchronisation of the matcher
        finally:
            torch.backends.cuda.matmul.allow_tf32 = tf32
        result = []
        for l in range(len(layer_outputs)):
            per_image = []
            for b in range(B):
                i, j = linear_sum_assignment(host[l, b, :, :counts[b]])
                per_image.append((torch.as_tensor(i, dtype=torch.int64), torch.as_tensor(j, dtype=torch.int64)))
            result.append(per_image)
        return result

    @torch.no_grad()
    def forward(self, outputs, targets, point_source=None):
        """Reference signature (matcher.py:160-179): one layer's predictions -> list of (index_i, index_j)."""
        return self.match_layers([outputs], targets, point_source)[0]

    def __repr__(self, _repr_indent=4):
        head = "Matcher " + self.__class__.__name__
        body = [f"cost_class: {self.cost_class}", f"cost_mask: {self.cost_mask}", f"cost_dice: {self.cost_dice}"]
        return "\n".join([head] + [" " * _repr_indent + line for line in body])

"""Point sampling for the training losses.

The reference imports ``point_sample`` and ``get_uncertain_point_coords_with_randomness`` from detectron2's PointRend
project (criterion.py:13-16, matcher.py:12) - third-party code that is not part of /root/reference (detectron2 0.6,
projects/PointRend/point_rend/point_features.py). Their published behaviour is restated here:

* ``point_sample``: bilinear ``grid_sample`` at normalised [0,1] x [0,1] (x, y) coordinates, zeros outside.
* importance sampling: draw ``oversample_ratio * P`` uniform points per mask, keep the ``importance_sample_ratio * P``
  most uncertain ones, fill up with fresh uniform points.

``PointSource`` is the one place random numbers are drawn, for ALL decoder layers of a step in three calls (the
reference draws per layer and, in the matcher, per image); tests inject recorded coordinates through it.
"""
import torch
import torch.nn.functional as F


def point_sample(input, point_coords, **kwargs):
    """input [N,C,H,W], point_coords [N,P,2] in [0,1]^2 (x, y) -> [N,C,P]."""
    add_dim = point_coords.dim() == 3
    if add_dim:
        point_coords = point_coords.unsqueeze(2)
    out = F.grid_sample(input, 2.0 * point_coords - 1.0, **kwargs)
    return out.squeeze(3) if add_dim else out


class PointSource:
    """Uniform [0,1) coordinates on ``device``; one draw covers every layer."""

    def matcher_points(self, layers, batch, num_points, device):
        """[layers, batch, P, 2]: one point set per image and layer (matcher.py:117)."""
        return torch.rand(layers, batch, num_points, 2, device=device)

    def oversampled_points(self, layers, num_masks, num_sampled, device):
        """[layers, N, oversample * P, 2] candidates of the importance sampling."""
        return torch.rand(layers, num_masks, num_sampled, 2, device=device)

    def random_points(self, layers, num_masks, num_random, device):
        """[layers, N, P - importance * P, 2] uniform fill-up points."""
        return torch.rand(layers, num_masks, num_random, 2, device=device)


def uncertain_point_coords(logits, candidates, fill, num_points, importance_sample_ratio):
    """Importance sampling of get_uncertain_point_coords_with_randomness with the random draws passed in.

    logits [R,1,h,w]; candidates [R,S,2]; fill [R,P - k,2] (k = int(importance_sample_ratio * P)); uncertainty is
    -|logit| (criterion.py:75-90). Returns [R,P,2]."""
    R = logits.shape[0]
    k = int(importance_sample_ratio * num_points)
    unc = -point_sample(logits, candidates, align_corners=False)[:, 0, :].abs()
    idx = torch.topk(unc, k=k, dim=1)[1]
    coords = torch.gather(candidates, 1, idx.unsqueeze(-1).expand(R, k, 2))
    if num_points - k > 0:
        coords = torch.cat([coords, fill], dim=1)
    return coords

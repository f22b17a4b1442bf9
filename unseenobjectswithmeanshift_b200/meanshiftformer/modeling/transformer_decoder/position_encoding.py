"""Sine positional embedding (reference: modeling/transformer_decoder/position_encoding.py:12-52).

With no padding mask (the only way the hot path calls it) the embedding is separable: channel
block [0, npf) depends on y only, block [npf, 2*npf) on x only. ``table`` builds the two small
tables once per (H, W) and the decoders add them where the reference materialises a
[B, 2*npf, H, W] tensor per forward (315 MB per image at 480x640).
"""
import math

import torch
from torch import nn


class PositionEmbeddingSine(nn.Module):
    def __init__(self, num_pos_feats=64, temperature=10000, normalize=False, scale=None):
        super().__init__()
        self.num_pos_feats = num_pos_feats
        self.temperature = temperature
        self.normalize = normalize
        if scale is not None and normalize is False:
            raise ValueError("normalize should be True if scale is passed")
        if scale is None:
            scale = 2 * math.pi
        self.scale = scale
        self._cache = {}

    def _axis_table(self, length, device):
        """[length, npf]: sin/cos interleaved features of the 1-based, normalised coordinate."""
        e = torch.arange(1, length + 1, dtype=torch.float32, device=device)  # cumsum of ones
        if self.normalize:
            e = e / (e[-1:] + 1e-6) * self.scale
        dim_t = torch.arange(self.num_pos_feats, dtype=torch.float32, device=device)
        dim_t = self.temperature ** (2 * torch.div(dim_t, 2, rounding_mode="floor") / self.num_pos_feats)
        p = e[:, None] / dim_t
        return torch.stack((p[:, 0::2].sin(), p[:, 1::2].cos()), dim=2).flatten(1)

    def tables(self, height, width, device):
        """(ty [H, npf], tx [W, npf]), cached per size and device."""
        key = (height, width, str(device))
        if key not in self._cache:
            self._cache[key] = (self._axis_table(height, device), self._axis_table(width, device))
        return self._cache[key]

    def table(self, height, width, device):
        """[H*W, 2*npf] rows in raster order: the same values as forward(...).flatten(2).T for one image."""
        key = ("full", height, width, str(device))
        if key in self._cache:
            return self._cache[key]
        ty, tx = self.tables(height, width, device)
        npf = self.num_pos_feats
        t = torch.cat((ty[:, None, :].expand(height, width, npf), tx[None, :, :].expand(height, width, npf)),
                      dim=2).reshape(height * width, 2 * npf)
        if t.numel() * 4 <= (64 << 20):  # keep small tables; the full-resolution one (315 MB) is rebuilt per call
            self._cache[key] = t
        return t

    def forward(self, x, mask=None):
        if mask is not None:  # padded batches: the reference's cumsum formulation
            not_mask = ~mask
            y_embed = not_mask.cumsum(1, dtype=torch.float32)
            x_embed = not_mask.cumsum(2, dtype=torch.float32)
            if self.normalize:
                y_embed = y_embed / (y_embed[:, -1:, :] + 1e-6) * self.scale
                x_embed = x_embed / (x_embed[:, :, -1:] + 1e-6) * self.scale
            dim_t = torch.arange(self.num_pos_feats, dtype=torch.float32, device=x.device)
            dim_t = self.temperature ** (2 * torch.div(dim_t, 2, rounding_mode="floor") / self.num_pos_feats)
            px, py = x_embed[..., None] / dim_t, y_embed[..., None] / dim_t
            px = torch.stack((px[..., 0::2].sin(), px[..., 1::2].cos()), dim=4).flatten(3)
            py = torch.stack((py[..., 0::2].sin(), py[..., 1::2].cos()), dim=4).flatten(3)
            return torch.cat((py, px), dim=3).permute(0, 3, 1, 2)
        B, _, H, W = x.shape
        t = self.table(H, W, x.device)
        return t.t().reshape(1, 2 * self.num_pos_feats, H, W).expand(B, -1, -1, -1)

    def __repr__(self, _repr_indent=4):
        head = "Positional encoding " + self.__class__.__name__
        body = [f"num_pos_feats: {self.num_pos_feats}", f"temperature: {self.temperature}",
                f"normalize: {self.normalize}", f"scale: {self.scale}"]
        return "\n".join([head] + [" " * _repr_indent + line for line in body])

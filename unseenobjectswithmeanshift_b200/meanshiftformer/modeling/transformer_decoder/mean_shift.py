"""Classical vMF mean-shift clustering - mirror of the reference's
modeling/transformer_decoder/mean_shift.py (= lib/utils/mean_shift.py) with the hill climbing in CUDA.

Same function names and argument meaning; cosine metric only (the one every UOIS config uses).
All tensors are CUDA fp32; ``seed_hill_climbing_ball`` additionally accepts a batch axis
(X [B,n,d], Z [B,m,d]) so a whole batch of images is one launch sequence.
"""
import numpy as np
import torch

from .... import ops


def _cosine_only(metric):
    if metric != "cosine":
        raise NotImplementedError("only metric='cosine' is implemented (EMBEDDING_METRIC of all UOIS configs)")


def seed_hill_climbing_ball(X, Z, kappa, max_iters=10, metric="cosine"):
    """Reference mean_shift.py:79-109:  repeat max_iters:  Z <- unit(exp(kappa Z X^T) X)."""
    _cosine_only(metric)
    return ops.mean_shift_hill_climb(X, Z, kappa, max_iters)


def connected_components(Z, epsilon, metric="cosine"):
    """Reference mean_shift.py:41-76 - a sequential sweep over the m (~100) converged seeds, run by one CTA
    on the device (ops.seed_connected_components). Returns a CPU LongTensor like the reference."""
    _cosine_only(metric)
    return ops.seed_connected_components(Z, epsilon)[0].cpu()


def mean_shift_with_seeds(X, Z, kappa, max_iters=10, metric="cosine", cfg_TRAIN_EMBEDDING_ALPHA=0.02):
    """Reference mean_shift.py:112-125."""
    Z = seed_hill_climbing_ball(X, Z, kappa, max_iters=max_iters, metric=metric)
    return connected_components(Z, 2 * cfg_TRAIN_EMBEDDING_ALPHA, metric=metric), Z


def select_smart_seeds(X, num_seeds, return_selected_indices=False, init_seeds=None, num_init_seeds=None,
                       metric="cosine", first_index=None):
    """Reference mean_shift.py:128-189 (farthest-point seeding): one cooperative CUDA launch that keeps a running
    nearest-seed distance per point instead of the reference's growing [n, i] matrix (same arg-max sequence).
    ``first_index`` (extension) fixes the first seed; default draws np.random.randint like :155.
    ``init_seeds`` / ``num_init_seeds`` (no caller in the reference) are not implemented."""
    _cosine_only(metric)
    if init_seeds is not None:
        raise NotImplementedError("init_seeds is not implemented (unused by every caller of the reference)")
    first = int(np.random.randint(0, X.shape[0])) if first_index is None else int(first_index)
    seeds, idx = ops.select_smart_seeds(X, num_seeds, [first])
    return (seeds, idx.cpu()) if return_selected_indices else (seeds,)


def mean_shift_smart_init_batched(X, kappa, num_seeds=100, max_iters=10, first_index=None, alpha=0.02):
    """mean_shift_smart_init (mean_shift.py:192-229) for a batch X [B,n,d]: seeding, hill climb, seed merging,
    assignment and relabelling all stay on the device (24 kernel launches + 2 memsets per 10-iteration call, no
    synchronisation). Returns (labels int64 [B,n], selected int64 [B,num_seeds], converged seeds [B,num_seeds,d])."""
    B, n = X.shape[0], X.shape[1]
    if first_index is None:
        first_index = np.random.randint(0, n, size=B)
    seeds, selected = ops.select_smart_seeds(X, num_seeds, first_index)
    Z = ops.mean_shift_hill_climb(X, seeds, kappa, max_iters)
    seed_labels, num = ops.seed_connected_components(Z, 2 * alpha)
    return ops.assign_clusters(X, Z, seed_labels, num), selected, Z


def mean_shift_smart_init(X, kappa, num_seeds=100, max_iters=10, metric="cosine", first_index=None):
    """Reference mean_shift.py:192-229: (cluster labels [n], indices of the selected seeds)."""
    _cosine_only(metric)
    labels, selected, _ = mean_shift_smart_init_batched(
        X.unsqueeze(0), kappa, num_seeds, max_iters, None if first_index is None else [int(first_index)])
    return labels[0], selected[0].cpu()


def clustering_features(features, num_seeds=100, kappa=20, max_iters=10, first_index=None):
    """Reference lib/fcn/test_dataset.py:43-59: features [B,C,H,W] (unit along C) -> (out_label float [B,H,W],
    list of the selected pixel indices per image); the per-image Python loop becomes one batched device pass."""
    B, C, H, W = features.shape
    X = features.reshape(B, C, H * W).transpose(1, 2).contiguous()
    labels, selected, _ = mean_shift_smart_init_batched(X, kappa, num_seeds, max_iters, first_index)
    return labels.view(B, H, W).float(), list(selected.cpu())

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
    """Reference mean_shift.py:41-76 - a sequential sweep over the m (~100) converged seeds.
    Host logic on a [m,m] distance matrix computed once on the device; returns a CPU LongTensor
    like the reference."""
    _cosine_only(metric)
    dist = (0.5 * (1 - Z @ Z.t())).cpu().numpy()
    n = dist.shape[0]
    labels = np.full(n, -1, dtype=np.int64)
    K = 0
    for i in range(n):
        if labels[i] != -1:
            continue
        member = dist[:, i] <= epsilon
        current = labels[member]
        if np.unique(current).shape[0] > 1:
            seen = current[current != -1]
            vals, counts = np.unique(seen, return_counts=True)
            lab = vals[np.argmax(counts)]
        else:
            lab = K
            K += 1
        labels[member] = lab
    return torch.from_numpy(labels)


def mean_shift_with_seeds(X, Z, kappa, max_iters=10, metric="cosine", cfg_TRAIN_EMBEDDING_ALPHA=0.02):
    """Reference mean_shift.py:112-125."""
    Z = seed_hill_climbing_ball(X, Z, kappa, max_iters=max_iters, metric=metric)
    return connected_components(Z, 2 * cfg_TRAIN_EMBEDDING_ALPHA, metric=metric), Z


def select_smart_seeds(X, num_seeds, return_selected_indices=False, init_seeds=None, num_init_seeds=None,
                       metric="cosine", first_index=None):
    """Reference mean_shift.py:128-189 (farthest-point seeding) with a running nearest-seed
    distance instead of the reference's growing [n, i] matrix (same arg-max sequence).
    ``first_index`` (extension) fixes the first seed; default draws np.random.randint like :155."""
    _cosine_only(metric)
    n = X.shape[0]
    idx = torch.full((num_seeds,), -1, dtype=torch.long)
    if init_seeds is None:
        seeds = torch.empty((num_seeds, X.shape[1]), device=X.device)
        first = int(np.random.randint(0, n)) if first_index is None else int(first_index)
        idx[0] = first
        seeds[0] = X[first]
        nearest = 0.5 * (1 - X @ X[first])
        chosen = 1
    else:
        seeds, chosen = init_seeds, num_init_seeds
        nearest = (0.5 * (1 - X @ seeds[:chosen].t())).min(dim=1)[0]
    for i in range(chosen, num_seeds):
        j = torch.argmax(nearest)
        idx[i] = j
        seeds[i] = X[j]
        nearest = torch.minimum(nearest, 0.5 * (1 - X @ X[j]))
    return (seeds, idx) if return_selected_indices else (seeds,)


def mean_shift_smart_init(X, kappa, num_seeds=100, max_iters=10, metric="cosine", first_index=None):
    """Reference mean_shift.py:192-229."""
    seeds, selected = select_smart_seeds(X, num_seeds, return_selected_indices=True, metric=metric,
                                         first_index=first_index)
    seed_labels, Z = mean_shift_with_seeds(X, seeds, kappa, max_iters=max_iters, metric=metric)
    closest = torch.argmin(0.5 * (1 - X @ Z.t()), dim=1)
    labels = seed_labels.to(X.device)[closest]
    num = len(torch.unique(seed_labels))
    count = torch.bincount(labels, minlength=num)[:num]
    big = int(torch.argmax(count))
    if big != 0:
        a, b = labels == 0, labels == big
        labels[a] = big
        labels[b] = 0
    return labels, selected

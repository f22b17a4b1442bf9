"""Oracle: classical von-Mises-Fisher mean-shift clustering. Test infrastructure only.

Follows MSMFormer/meanshiftformer/modeling/transformer_decoder/mean_shift.py (twin of
lib/utils/mean_shift.py, which differs only in where alpha comes from, :9,112,123).
Cosine metric only - the one the UOIS configs use (experiments/cfgs/*: EMBEDDING_METRIC cosine).
"""
import numpy as np
import torch
import torch.nn.functional as F


def seed_hill_climbing_ball(X, Z, kappa, max_iters=10):
    """mean_shift.py:79-109 with ball_kernel :11-27 (cosine):
    repeat  W = exp(kappa * Z X^T);  Z = rows of (W X) rescaled to unit length."""
    for _ in range(max_iters):
        W = torch.exp(kappa * torch.mm(Z, X.t()))
        Z = F.normalize(torch.mm(W, X), p=2, dim=1)
    return Z


def connected_components(Z, epsilon):
    """mean_shift.py:41-76. Sequential sweep over seeds: every seed within cosine distance
    epsilon of seed i joins i's component; if some of them already carry labels, the most
    frequent existing label (ties -> smallest) is reused, otherwise a fresh label is opened."""
    n = Z.shape[0]
    labels = np.full(n, -1, dtype=np.int64)
    K = 0
    Zn = Z.detach()
    for i in range(n):
        if labels[i] != -1:
            continue
        dist = 0.5 * (1 - torch.mm(Zn, Zn[i:i + 1].t()))[:, 0]
        member = (dist <= epsilon).numpy()
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


def mean_shift_with_seeds(X, Z, kappa, max_iters=10, alpha=0.02):
    """mean_shift.py:112-125; epsilon = 2*alpha (cfg.TRAIN.EMBEDDING_ALPHA, lib/fcn/config.py:255)."""
    Z = seed_hill_climbing_ball(X, Z, kappa, max_iters)
    return connected_components(Z, 2 * alpha), Z


def select_smart_seeds(X, num_seeds, first_index):
    """mean_shift.py:128-189 (farthest-point seeding). ``first_index`` replaces the reference's
    np.random.randint(0, n) draw (:155) so that the caller controls the randomness.
    Returns (seeds [m,d], indices int64 [m])."""
    n = X.shape[0]
    idx = torch.empty(num_seeds, dtype=torch.long)
    idx[0] = first_index
    nearest = 0.5 * (1 - torch.mm(X, X[first_index].unsqueeze(1))[:, 0])
    for i in range(1, num_seeds):
        j = torch.argmax(nearest)
        idx[i] = j
        nearest = torch.minimum(nearest, 0.5 * (1 - torch.mm(X, X[j].unsqueeze(1))[:, 0]))
    return X[idx].clone(), idx


def mean_shift_smart_init(X, kappa, num_seeds=100, max_iters=10, first_index=0, alpha=0.02):
    """mean_shift.py:192-229: seed, climb, merge seeds, label every point by its closest seed,
    then swap labels so that the most populous cluster is 0."""
    seeds, idx = select_smart_seeds(X, num_seeds, first_index)
    seed_labels, Z = mean_shift_with_seeds(X, seeds, kappa, max_iters, alpha)
    closest = torch.argmin(0.5 * (1 - torch.mm(X, Z.t())), dim=1)
    labels = seed_labels[closest]
    num = len(torch.unique(seed_labels))
    count = torch.stack([(labels == i).sum() for i in range(num)])
    big = int(torch.argmax(count))
    if big != 0:
        a, b = labels == 0, labels == big
        labels[a] = big
        labels[b] = 0
    return labels, idx

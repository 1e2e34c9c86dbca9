"""Training-side nearest-neighbour utilities on the distance kernel.

The reference's trainer builds two small Euclidean distance matrices with ``torch.cdist`` per
epoch -- the minimum distance between proxies of different identities
(``train_encodersKIT.py:146-153``) and the farthest-point ("triangulation") proxy selection
(``train_encodersKIT.py:252-284``).  Both are the a2' metric of this library
(``DALI_METRIC_EUCLIDEAN``); the few reductions around it stay torch ops on the device, in the
reference's own order.  Same function names, arguments and return values.
"""
from __future__ import annotations

import numpy as np
import torch

from . import metrics

__all__ = ["min_negative_distance", "selectProxiesByTriagulation"]


def _cdist(X, precision):
    X = X if isinstance(X, torch.Tensor) else torch.as_tensor(np.asarray(X))
    if not X.is_cuda:
        X = X.cuda()
    return metrics.compute_distance_matrix(X.float().contiguous(), X.float().contiguous(), "euclidean",
                                           precision=precision, normalize=False)


def min_negative_distance(all_proxies, proxies_labels, precision="tf32c"):
    """``train_encodersKIT.py:146-153``: smallest Euclidean distance between two proxies with
    different labels (same-label pairs are lifted to the global maximum first)."""
    d = _cdist(all_proxies, precision)
    lab = torch.as_tensor(np.asarray(proxies_labels).astype(np.int64), device=d.device)
    mask = (lab[:, None] == lab[None, :]).to(d.dtype)
    d = mask * torch.max(d) + (1 - mask) * d
    return torch.min(d).item()


def selectProxiesByTriagulation(X, num_proxies=5, precision="tf32c"):
    """``train_encodersKIT.py:252-284``: start from a random sample (``np.random.choice``, as the
    reference), then repeatedly add the sample farthest from the chosen set.  Returns
    ``(proxies LongTensor, max distance between proxies)``."""
    dist = _cdist(X, precision)
    n = dist.shape[0]
    cumulative_vector = torch.ones(n, device=dist.device) * torch.max(dist)
    proxies = [int(np.random.choice(n))]
    num_proxies = min(num_proxies, n)
    i = 0
    for _ in range(num_proxies - 1):
        sample_idx = proxies[i]
        cumulative_vector = torch.minimum(cumulative_vector, dist[sample_idx])
        furthest_idx = torch.argsort(cumulative_vector, stable=True)[-1]
        proxies.append(int(furthest_idx.item()))
        i += 1
    proxies = torch.tensor(proxies, dtype=torch.long)
    sel = proxies.to(dist.device)
    max_dist_between_proxies = torch.max(dist[sel, :][:, sel]).item()
    return proxies, max_dist_between_proxies

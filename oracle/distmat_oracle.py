"""CPU ORACLE (test infrastructure, NOT product code) for the distance / fusion /
top-k arithmetic.  Every function is the reference's own expression, executed with
CPU torch / numpy exactly as the reference writes it (these lines need no
third-party code, so this part of the oracle is the reference itself run here).
"""
from __future__ import annotations

import numpy as np
import torch


def l2_normalize(x: torch.Tensor) -> torch.Tensor:
    """``x/torch.norm(x, dim=1, keepdim=True)`` -- validateModels.py:41-42,
    evaluate.py:285-286 (no eps: a zero row becomes NaN)."""
    return x / torch.norm(x, dim=1, keepdim=True)


def cosine_distmat(qf: torch.Tensor, gf: torch.Tensor, normalize: bool = True) -> torch.Tensor:
    """``1.0 - torch.mm(q, g.T)`` -- validateModels.py:47, evaluate.py:291."""
    if normalize:
        qf, gf = l2_normalize(qf), l2_normalize(gf)
    return 1.0 - torch.mm(qf, gf.T)


def sqeuclidean_distmat(qf: torch.Tensor, gf: torch.Tensor) -> torch.Tensor:
    """``torchreid.metrics.compute_distance_matrix(q, g, "euclidean")`` (commented call at
    validateModels.py:44, evaluate.py:288) [upstream-recall]: squared Euclidean,
    ``|q|^2 + |g|^2`` then ``addmm_(q, g.T, beta=1, alpha=-2)``; no sqrt, no clamp."""
    m, n = qf.size(0), gf.size(0)
    mat1 = torch.pow(qf, 2).sum(dim=1, keepdim=True).expand(m, n)
    mat2 = torch.pow(gf, 2).sum(dim=1, keepdim=True).expand(n, m).t()
    distmat = mat1 + mat2
    distmat.addmm_(qf, gf.t(), beta=1, alpha=-2)
    return distmat


def euclidean_distmat(qf: torch.Tensor, gf: torch.Tensor) -> torch.Tensor:
    """``torch.cdist(q, g, p=2.0)`` -- commented call at validateModels.py:45, evaluate.py:289."""
    return torch.cdist(qf, gf, p=2.0)


def fuse_mean(distmats):
    """``(d1+d2)/2`` evaluate_ensembled_models.py:313, evaluateCleanATModels.py:127;
    ``(d_bb+d_h1+d_h2)/3`` evaluate.py:278.  numpy fp32, left-to-right."""
    acc = np.asarray(distmats[0])
    for d in distmats[1:]:
        acc = acc + np.asarray(d)
    return acc / len(distmats)


def magnitude_weights(q_mag: torch.Tensor, g_mag: torch.Tensor) -> torch.Tensor:
    """``torch.maximum(q_mag.repeat(1,G), g_mag.T.repeat(Q,1))`` --
    evaluateCleanATModels.py:154-155 (``q_mag`` is [Q,1], ``g_mag`` is [G,1])."""
    return torch.maximum(q_mag.repeat(1, g_mag.shape[0]), g_mag.T.repeat(q_mag.shape[0], 1))


def fuse_weighted(w_list, d_list):
    """``(w_c*d_c + w_d*d_d)/(w_c + w_d)`` -- evaluateCleanATModels.py:157 (torch fp32;
    the distmats are numpy arrays multiplied into torch tensors)."""
    num = w_list[0] * torch.as_tensor(np.asarray(d_list[0]))
    den = w_list[0]
    for w, d in zip(w_list[1:], d_list[1:]):
        num = num + w * torch.as_tensor(np.asarray(d))
        den = den + w
    return num / den


def briar_topk(distmat: torch.Tensor, k: int = 20) -> torch.Tensor:
    """``torch.argsort(distmat, dim=1)[:, :20]`` -- validateModels.py:93, with the
    canonical (stable) tie order."""
    return torch.argsort(distmat, dim=1, stable=True)[:, :k]

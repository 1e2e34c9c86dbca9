"""Verification metric of the reference's dead ROC branch (SURVEY 8f row N3,
evaluateCleanATModels.py:276-292).

NOT part of the hot path and NOT a hand-written kernel: the ROC over all Q x G pairs is the
stock-library formulation on the device (``torch.sort`` over the scores + cumulative sums).  It is
kept in its own module so that ``metrics.py`` stays "ctypes only, no arithmetic".
"""
from __future__ import annotations

import numpy as np

from .metrics import canonicalize_labels

__all__ = ["roc_curve_pairs"]


def roc_curve_pairs(distmat, q_pids, g_pids, device=None):
    """ROC over all Q x G pairs, as the (dead) verification branch of the reference computes it
    (evaluateCleanATModels.py:276-292): label = same identity, score = ``1.0 - distmat/2.0``,
    ``sklearn.metrics.roc_curve(labels, scores, pos_label=1)`` -> ``(fpr, tpr, thresholds)``.

    SURVEY 8f row N3.  This is NOT a hand-written kernel: it is the stock-library formulation on the
    device (``torch.sort`` over the Q*G scores + cumulative sums), kept out of the hot path.  The
    result equals scikit-learn's (1.9: first threshold ``inf``, intermediate collinear points
    dropped) element for element; 53.6 M pairs take tens of milliseconds instead of ~10 s."""
    import torch
    d = distmat if isinstance(distmat, torch.Tensor) else torch.as_tensor(np.asarray(distmat, dtype=np.float32))
    if not d.is_cuda:
        d = d.cuda(device if device is not None else torch.cuda.current_device())
    d = d.float()
    qp, gp = canonicalize_labels(q_pids, g_pids)
    qpt = torch.as_tensor(qp, device=d.device)
    gpt = torch.as_tensor(gp, device=d.device)
    score = 1.0 - d.reshape(-1) / 2.0
    y = (qpt[:, None] == gpt[None, :]).reshape(-1)
    n = score.numel()
    order = torch.argsort(score, stable=True).flip(0)  # np.argsort(kind="mergesort")[::-1]
    ys = score[order]
    yt = y[order].to(torch.float64)
    distinct = torch.nonzero(ys[1:] != ys[:-1]).reshape(-1)
    idx = torch.cat([distinct, torch.tensor([n - 1], device=d.device)])
    tps = torch.cumsum(yt, 0)[idx]
    fps = 1.0 + idx.to(torch.float64) - tps
    thr = ys[idx]
    if idx.numel() > 2:  # drop_intermediate=True
        d2f = torch.diff(fps, n=2) != 0
        d2t = torch.diff(tps, n=2) != 0
        keep = torch.cat([torch.tensor([True], device=d.device), d2f | d2t,
                          torch.tensor([True], device=d.device)])
        tps, fps, thr = tps[keep], fps[keep], thr[keep]
    z = torch.zeros(1, dtype=torch.float64, device=d.device)
    tps = torch.cat([z, tps])
    fps = torch.cat([z, fps])
    thr = torch.cat([torch.full((1,), float("inf"), dtype=thr.dtype, device=d.device), thr])
    fpr = fps / fps[-1]
    tpr = tps / tps[-1]
    return fpr.cpu().numpy(), tpr.cpu().numpy(), thr.cpu().numpy()

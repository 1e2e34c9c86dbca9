"""Verification metric of the reference's dead ROC branch (SURVEY 8f row N3,
evaluateCleanATModels.py:276-292).

``roc_curve_binned`` is the kernel path (``csrc/roc.cu`` through ``dali_roc_hist_f32``): one
streaming pass over the distance matrix histograms the scores of the two classes; the suffix sums
are exact points of the ROC at ``bins`` thresholds -- no sort, no Q x G temporaries.
``roc_curve_pairs`` is the stock-library formulation (``torch.sort`` + cumulative sums) that
reproduces scikit-learn's output element for element; it is the checker of the kernel path, not a
kernel.
"""
from __future__ import annotations

import numpy as np

from .metrics import canonicalize_labels

__all__ = ["roc_curve_pairs", "roc_curve_binned"]


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


def roc_curve_binned(distmat, q_pids, g_pids, bins=65536, lo=0.0, hi=1.0):
    """ROC of evaluateCleanATModels.py:276-292 (label = same identity, score = ``1.0 - distmat/2.0``)
    at ``bins`` thresholds, by the histogram kernel (``dali_roc_hist_f32``).

    Returns ``(fpr, tpr, thresholds)`` with ``bins + 1`` points, thresholds descending like
    scikit-learn's: point 0 is ``(0, 0, inf)``, point ``i`` counts the scores in bins ``>= bins - i``.
    Every point lies exactly on the curve ``sklearn.metrics.roc_curve`` returns (its counts are the
    exact numbers of positives / negatives at that threshold); ``thresholds[i]`` is the lower edge
    of the bin, ``lo + (bins - i) * (hi - lo) / bins``.  Scores outside ``[lo, hi]`` fall into the
    end bins (cosine distances in [0, 2] give scores in [0, 1])."""
    import ctypes

    from . import _lib
    from ._lib import as_matrix, c_vp, p_i32
    d = as_matrix(distmat, np.float32, "distmat")
    Q, G = d.shape
    qp, gp = canonicalize_labels(q_pids, g_pids)
    if qp.shape[0] != Q or gp.shape[0] != G:
        raise ValueError("label arrays do not match the distance matrix shape")
    ctx = _lib.get_ctx(d.device)
    ctx.attach_torch_stream()
    pos = np.zeros(bins, dtype=np.uint64)
    neg = np.zeros(bins, dtype=np.uint64)
    ctx.check(ctx.lib.dali_roc_hist_f32(ctx.h, c_vp(d.ptr), Q, G, max(d.ld, 1), p_i32(qp), p_i32(gp), int(bins),
                                        ctypes.c_float(lo), ctypes.c_float(hi), c_vp(pos.ctypes.data),
                                        c_vp(neg.ctypes.data)))
    tps = np.concatenate([[0.0], np.cumsum(pos[::-1].astype(np.float64))])
    fps = np.concatenate([[0.0], np.cumsum(neg[::-1].astype(np.float64))])
    thr = np.concatenate([[np.inf], lo + np.arange(bins - 1, -1, -1, dtype=np.float64) * ((hi - lo) / bins)])
    return fps / max(fps[-1], 1.0), tps / max(tps[-1], 1.0), thr

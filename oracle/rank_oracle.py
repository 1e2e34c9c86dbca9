"""CPU ORACLE (test infrastructure, NOT product code) for the rank / CMC / mAP stage.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.  The product path
(``daliid_b200``) never does; it fails loudly when the CUDA library is missing.

What it restates
----------------
The reference calls a third-party evaluator at
``validateModels.py:68-69``, ``evaluate.py:312-313``,
``evaluate_ensembled_models.py:324-325`` and ``evaluateCleanATModels.py:266-267``::

    cmc, mAP = torchreid.metrics.evaluate_rank(distmat, q_pids, g_pids,
                                               q_camids, g_camids,
                                               use_metric_cuhk03=False)

``torchreid`` (PyPI ``torchreid`` / KaiyangZhou/deep-person-reid, version NOT
pinned by the reference: it ships no requirements file) is absent from
``/root/reference`` and from this image, so the published algorithm of
``torchreid/metrics/rank.py::eval_market1501`` and
``torchreid/metrics/rank_cylib/rank_cy.pyx::eval_market1501_cy`` is restated
here from the upstream's public semantics.

**PARITY UNPINNED**: the reference holds no golden vector, known-answer test or
fixture at this boundary, and the third-party module cannot
be executed here.  The oracle is therefore pinned only by (a) hand-derived
known-answer cases, (b) agreement between its two independently written
accumulation variants, (c) ``sklearn.metrics.average_precision_score`` on
tie-free inputs, (d) the C restatement in ``rank_oracle.c``.

Canonical tie order
-------------------
Upstream uses ``np.argsort(distmat, axis=1)`` whose default kind is not stable
(platform dependent).  The oracle *defines* the canonical order as the stable
one: distance ascending, gallery index ascending, NaN last (numpy's order),
``-0.0 == +0.0``.  ``tie='numpy_default'`` reproduces whatever this box's
numpy does (used for the timed CPU baseline only).

Two accumulation variants (both upstream behaviours)
----------------------------------------------------
``cy_f32``  what an installed torchreid runs when its Cython extension is built
            (the reference probes for it at ``validateModels.py:16-24``):
            C ``float`` state, AP accumulated sequentially in rank order with
            each term formed in double and the running sum stored back to
            float; mAP a sequential float sum in query order.
``py_f64``  the pure-Python fallback: float64 terms, ``ndarray.sum()``
            (numpy pairwise), ``np.mean`` over valid queries.
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "canonicalize_labels",
    "stable_argsort",
    "evaluate_rank",
    "eval_market1501_py_f64",
    "eval_market1501_cy_f32",
    "rank_details",
    "briar_rank_hits",
]

NO_VALID_MSG = "Error: all query identities do not appear in gallery"


def canonicalize_labels(q, g):
    """Map str/int label columns to dense int64 ids (equality is all that matters).

    The reference's label columns are *strings* (``datasetUtils.py:15-17`` builds
    ``np.array([[path, pid, camid, kind], ...])``), and ``queries[:,1]`` etc. are
    handed straight to the evaluator (``validateModels.py:68``).
    """
    q = np.asarray(q).reshape(-1)
    g = np.asarray(g).reshape(-1)
    both = np.concatenate([q, g])
    _, inv = np.unique(both, return_inverse=True)
    inv = inv.astype(np.int64)
    return inv[: q.shape[0]], inv[q.shape[0]:]


def stable_argsort(distmat):
    """Canonical order: distance asc, index asc, NaN last (== numpy kind='stable')."""
    return np.argsort(np.asarray(distmat), axis=1, kind="stable")


def _argsort(distmat, tie):
    if tie == "stable":
        return stable_argsort(distmat)
    if tie == "numpy_default":
        return np.argsort(distmat, axis=1)  # upstream's literal call
    raise ValueError(tie)


def _prep(distmat, q_pids, g_pids, q_camids, g_camids, max_rank):
    distmat = np.asarray(distmat)
    if distmat.ndim != 2:
        raise ValueError("distmat must be 2-D")
    num_q, num_g = distmat.shape
    q_pids, g_pids = canonicalize_labels(q_pids, g_pids)
    q_camids, g_camids = canonicalize_labels(q_camids, g_camids)
    if q_pids.shape[0] != num_q or g_pids.shape[0] != num_g:
        raise ValueError("label / distmat shape mismatch")
    if num_g < max_rank:
        # upstream prints: 'Note: number of gallery samples is quite small, got {}'
        max_rank = num_g
    return distmat, q_pids, g_pids, q_camids, g_camids, num_q, num_g, max_rank


def eval_market1501_py_f64(distmat, q_pids, g_pids, q_camids, g_camids, max_rank=50,
                           tie="stable", return_details=False):
    """Pure-Python upstream variant (``rank.py::eval_market1501``), float64 AP."""
    distmat, q_pids, g_pids, q_camids, g_camids, num_q, num_g, max_rank = _prep(
        distmat, q_pids, g_pids, q_camids, g_camids, max_rank)
    indices = _argsort(distmat, tie)
    matches = (g_pids[indices] == q_pids[:, np.newaxis]).astype(np.int32)

    all_cmc, all_AP = [], []
    ap_full = np.full(num_q, np.nan, dtype=np.float64)
    first_rank = np.full(num_q, -1, dtype=np.int64)
    num_valid_q = 0.0
    for q_idx in range(num_q):
        q_pid, q_camid = q_pids[q_idx], q_camids[q_idx]
        order = indices[q_idx]
        remove = (g_pids[order] == q_pid) & (g_camids[order] == q_camid)
        keep = np.invert(remove)
        raw_cmc = matches[q_idx][keep]
        if not np.any(raw_cmc):
            continue
        cmc = raw_cmc.cumsum()
        cmc[cmc > 1] = 1
        row = cmc[:max_rank]
        if row.shape[0] < max_rank:
            # kept gallery shorter than max_rank (undefined upstream: ragged list /
            # stale Cython buffer).  Defined here: the row saturates at its last value.
            row = np.concatenate([row, np.full(max_rank - row.shape[0], row[-1], row.dtype)])
        all_cmc.append(row)
        num_valid_q += 1.0
        num_rel = raw_cmc.sum()
        tmp_cmc = raw_cmc.cumsum()
        tmp_cmc = [x / (i + 1.0) for i, x in enumerate(tmp_cmc)]
        tmp_cmc = np.asarray(tmp_cmc) * raw_cmc
        AP = tmp_cmc.sum() / num_rel
        all_AP.append(AP)
        ap_full[q_idx] = AP
        first_rank[q_idx] = int(np.argmax(raw_cmc)) + 1
    assert num_valid_q > 0, NO_VALID_MSG
    all_cmc = np.asarray(all_cmc).astype(np.float32)
    all_cmc = all_cmc.sum(0) / num_valid_q
    mAP = np.mean(all_AP)
    if return_details:
        return all_cmc.astype(np.float32), float(mAP), ap_full, first_rank
    return all_cmc.astype(np.float32), float(mAP)


def eval_market1501_cy_f32(distmat, q_pids, g_pids, q_camids, g_camids, max_rank=50,
                           tie="stable", return_details=False):
    """Cython upstream variant (``rank_cy.pyx::eval_market1501_cy``), C-float state.

    Only the positions where ``raw_cmc == 1`` change the running sums (a zero
    term leaves a float unchanged), so the loop below visits positives only;
    the arithmetic per visited position is the upstream's:
    ``tmp_cmc_sum(float) += (tmp_cmc[g](float) / (g + 1.)(double)) * raw_cmc[g](float)``.
    """
    distmat, q_pids, g_pids, q_camids, g_camids, num_q, num_g, max_rank = _prep(
        distmat, q_pids, g_pids, q_camids, g_camids, max_rank)
    distmat = distmat.astype(np.float32, copy=False)  # upstream casts on entry
    indices = _argsort(distmat, tie)
    f32 = np.float32

    cmc_cnt = np.zeros(max_rank, dtype=np.int64)
    all_AP = np.zeros(num_q, dtype=np.float32)
    first_rank = np.full(num_q, -1, dtype=np.int64)
    num_valid_q = 0
    for q_idx in range(num_q):
        order = indices[q_idx]
        gp = g_pids[order]
        match = gp == q_pids[q_idx]
        keep = ~(match & (g_camids[order] == q_camids[q_idx]))
        raw = match[keep]
        pos = np.flatnonzero(raw)  # 0-based kept positions of the positives
        if pos.size == 0:
            continue
        num_valid_q += 1
        fr = int(pos[0]) + 1
        first_rank[q_idx] = fr
        if fr <= max_rank:
            cmc_cnt[fr - 1:] += 1
        s = f32(0.0)
        for k, p in enumerate(pos, start=1):
            term = float(f32(k)) / (float(p) + 1.0)  # double
            s = f32(float(s) + term)  # stored back to C float
        num_rel = f32(pos.size)
        all_AP[q_idx] = f32(s / num_rel)
    assert num_valid_q > 0, NO_VALID_MSG
    nv = f32(num_valid_q)
    cmc = (cmc_cnt.astype(np.float32) / nv).astype(np.float32)
    m = f32(0.0)
    for q_idx in range(num_q):
        m = f32(m + all_AP[q_idx])
    m = f32(m / nv)
    if return_details:
        ap_full = all_AP.astype(np.float64)
        ap_full[first_rank < 0] = np.nan
        return cmc, float(m), ap_full, first_rank
    return cmc, float(m)


def evaluate_rank(distmat, q_pids, g_pids, q_camids, g_camids, max_rank=50,
                  use_metric_cuhk03=False, use_cython=True, accum=None, tie="stable"):
    """Signature of ``torchreid.metrics.evaluate_rank`` [upstream-recall]."""
    if use_metric_cuhk03:
        raise NotImplementedError("cuhk03 metric is never used by the reference")
    if accum is None:
        accum = "cy_f32" if use_cython else "py_f64"
    fn = {"cy_f32": eval_market1501_cy_f32, "py_f64": eval_market1501_py_f64}[accum]
    return fn(distmat, q_pids, g_pids, q_camids, g_camids, max_rank, tie=tie)


def rank_details(distmat, q_pids, g_pids, q_camids, g_camids):
    """Per query: sorted 1-based kept ranks of every valid positive (stable order)."""
    distmat, q_pids, g_pids, q_camids, g_camids, num_q, num_g, _ = _prep(
        distmat, q_pids, g_pids, q_camids, g_camids, 50)
    indices = stable_argsort(distmat)
    out = []
    for q in range(num_q):
        order = indices[q]
        match = g_pids[order] == q_pids[q]
        keep = ~(match & (g_camids[order] == q_camids[q]))
        out.append(np.flatnonzero(match[keep]) + 1)
    return out


def briar_rank_hits(distmat, q_pids, g_pids, ranks=(1, 5, 10, 20), topk=20):
    """``validateBRIAR.calculateMetrics`` (``validateModels.py:84-105``) with the
    canonical stable order: closed-set top-k identification, no junk mask."""
    distmat = np.asarray(distmat)
    q_pids, g_pids = canonicalize_labels(q_pids, g_pids)
    ranked_idx = stable_argsort(distmat)[:, :topk]
    predicted = g_pids[ranked_idx]
    matching = q_pids.reshape(-1, 1) == predicted
    return [float(np.mean(np.sum(matching[:, :r], axis=1) > 0)) for r in ranks], ranked_idx

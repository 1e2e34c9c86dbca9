"""Mirrors of the evaluation arithmetic of the reference's stand-alone scripts
(``evaluate.py``, ``evaluate_ensembled_models.py``, ``evaluateCleanATModels.py``), with the
same function names and printed text.  The scripts' ``main()`` plumbing (datasets, model
factories, yacs config) is out of scope (it is not even importable as
shipped); what is kept is what those mains compute once features exist.
"""
from __future__ import annotations

import numpy as np

from . import metrics

RANKS = [1, 5, 10, 20]


def calculate_metrics(distmat, queries, gallery):
    """``calculate_metrics(distmat, queries, gallery)`` -- evaluate.py:305-320,
    evaluate_ensembled_models.py:317-332.  Prints, returns None (like the reference);
    the values are kept on ``calculate_metrics.last`` for callers that want them."""

    print('Computing CMC and mAP ...')

    cmc, mAP = metrics.evaluate_rank(distmat, queries[:, 1], gallery[:, 1],
                                     queries[:, 2], gallery[:, 2], use_metric_cuhk03=False)

    print('** Results **')
    print('mAP: {:.2%}'.format(mAP))
    print('CMC curve')
    for r in RANKS:
        print('Rank-{:<3}: {:.2%}'.format(r, cmc[r - 1]))
    calculate_metrics.last = (cmc, mAP)


def calculateMetrics(queries_images, gallery_images, distmat, pooling=None, version=None):
    """``calculateMetrics(queries_images, gallery_images, distmat, pooling=None, version=None)``
    -- evaluateCleanATModels.py:259-292.  With ``pooling`` set, the ROC over all pairs is computed
    and saved as ``FPR_<version>.npy``, ``TPR_<version>.npy``, ``Thresholds_<version>.npy`` exactly
    as the reference's (never exercised) branch does (276-292)."""
    calculate_metrics(distmat, queries_images, gallery_images)
    calculateMetrics.last = calculate_metrics.last
    if pooling:
        print((queries_images.shape[0], gallery_images.shape[0]), (queries_images.shape[0], gallery_images.shape[0]))
        from .verification import roc_curve_pairs
        fpr, tpr, thresholds = roc_curve_pairs(distmat, queries_images[:, 1], gallery_images[:, 1])
        np.save("FPR_%s" % version, fpr)
        np.save("TPR_%s" % version, tpr)
        np.save("Thresholds_%s" % version, thresholds)
        print("ROC Curve calculated!")


def evaluate_single_model(queries_fvs, gallery_fvs, queries, gallery, precision=metrics.DEFAULT_PRECISION):
    """Single-model branch of evaluate.py (282-302): normalise, ``1 - q.g``, metrics."""
    distmat = metrics.compute_distance_matrix(queries_fvs, gallery_fvs, "cosine", precision=precision)
    calculate_metrics(distmat, queries, gallery)
    return distmat


def evaluate_multiple_output(q_feats, g_feats, queries, gallery, precision=metrics.DEFAULT_PRECISION):
    """3-exit branch of evaluate.py (244-279): one distmat per exit, metrics per exit, then the
    mean ``(d_backbone + d_head01 + d_head02)/3`` (278) and its metrics (279)."""
    # the three contractions add their tiles to the running mean in their epilogues: no fusion pass
    distmats, distmat_ensemble = metrics.ensemble_distance_matrices(q_feats, g_feats, "cosine",
                                                                    precision=precision)
    for d in distmats:
        calculate_metrics(d, queries, gallery)
    calculate_metrics(distmat_ensemble, queries, gallery)
    return distmats, distmat_ensemble


def evaluate_ensembled_models(q_feats01, g_feats01, q_feats02, g_feats02, queries, gallery,
                              precision=metrics.DEFAULT_PRECISION):
    """evaluate_ensembled_models.py:274-314: two models, ``(distmat01+distmat02)/2``."""
    (distmat01, distmat02), distmat_ensemble = metrics.ensemble_distance_matrices(
        [q_feats01, q_feats02], [g_feats01, g_feats02], "cosine", precision=precision)
    calculate_metrics(distmat01, queries, gallery)
    calculate_metrics(distmat02, queries, gallery)
    calculate_metrics(distmat_ensemble, queries, gallery)
    return distmat01, distmat02, distmat_ensemble


def getWeightsByMagnitude(fvs):
    """Arithmetic of ``getWeightsByMagnitude`` (evaluateCleanATModels.py:249-256) once the
    pooled features exist: returns ``(magnitudes [N], fvs/magnitudes)``."""
    normed, norms = metrics.normalize(fvs, return_norms=True)
    return norms, normed


def magnitude_weighted_fusion(clean_distmat, distortion_distmat, q_mag_clean, g_mag_clean,
                              q_mag_distortion, g_mag_distortion):
    """``(w_c*d_c + w_d*d_d)/(w_c + w_d)`` with ``w = max(|q| repeated, |g|^T repeated)`` --
    evaluateCleanATModels.py:154-157 (GAP), 193-196 (GMP), 230-233 (both)."""
    return metrics.fuse_distmats([clean_distmat, distortion_distmat],
                                 q_weights=[q_mag_clean, q_mag_distortion],
                                 g_weights=[g_mag_clean, g_mag_distortion])


def evaluate_clean_at_models(q_clean, g_clean, q_dist, g_dist, queries_images, gallery_images,
                             magnitudes=None, precision=metrics.DEFAULT_PRECISION):
    """evaluateCleanATModels.py:103-160 from features: concatenated-feature distmat, per-model
    distmats, simple mean, and (when ``magnitudes`` = (q_mag_c, g_mag_c, q_mag_d, g_mag_d) is
    given) the magnitude-weighted ensemble."""
    import torch
    cat = (lambda a, b: torch.cat((torch.as_tensor(a), torch.as_tensor(b)), dim=1))
    concatenated = metrics.compute_distance_matrix(cat(q_clean, q_dist), cat(g_clean, g_dist),
                                                   "cosine", precision=precision)
    calculateMetrics(queries_images, gallery_images, concatenated)
    clean = metrics.compute_distance_matrix(q_clean, g_clean, "cosine", precision=precision)
    distortion = metrics.compute_distance_matrix(q_dist, g_dist, "cosine", precision=precision)
    simple = metrics.fuse_distmats([clean, distortion])
    calculateMetrics(queries_images, gallery_images, clean)
    calculateMetrics(queries_images, gallery_images, distortion)
    calculateMetrics(queries_images, gallery_images, simple)
    out = {"concatenated": concatenated, "clean": clean, "distortion": distortion, "mean": simple}
    if magnitudes is not None:
        ens = magnitude_weighted_fusion(clean, distortion, *magnitudes)
        calculateMetrics(queries_images, gallery_images, ens)
        out["weighted"] = ens
    return out


class Meta_Recognition(object):
    """``Meta_Recognition`` of evaluate.py:583-627 / evaluate_ensembled_models.py:593-637: the
    Weibull fitting (``libmr``) and the fusion run on the GPU (``csrc/mrfuse.cu``)."""

    def metarec(self, scorematrix, topk, use_columns=True, killscale=1):
        """Per-score weights ``[Q,G]`` fp64 (evaluate.py:587-608)."""
        _, det = metrics.mrfuse([scorematrix], topk, use_columns, killscale, return_details=True)
        return det["weights"][0]

    def mrfuse(self, scores01, scores02, scores03):
        """evaluate.py:610-627: ``metarec(., 20, use_columns=False)`` per model, weighted mean;
        numpy fp64 ``[Q,G]`` like the reference's ``scores.numpy()`` (the reference also prints
        three 5x5 weight corners, kept)."""
        fused, det = metrics.mrfuse([scores01, scores02, scores03], 20, False, 1, return_details=True)
        for w in det["weights"]:
            print(w[:5, :5])
        return fused.cpu().numpy() if hasattr(fused, "cpu") else fused

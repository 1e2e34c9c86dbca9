"""CPU ORACLE (test infrastructure, NOT product code) for k-reciprocal re-ranking
(SURVEY 8f row N1).

Only ``tests/`` may import this module; the product path (``daliid_b200``) never does.

What it restates
----------------
Every distance-matrix site of the reference carries a commented-out hook::

    #if rerank:
    #  distmat_qq = torchreid.metrics.compute_distance_matrix(q, q, metric="euclidean")
    #  distmat_gg = torchreid.metrics.compute_distance_matrix(g, g, metric="euclidean")
    #  distmat = torchreid.utils.re_ranking(distmat, distmat_qq, distmat_gg)

(``validateModels.py:49-53``, ``evaluate.py:294-298``, ``evaluate_ensembled_models.py:284-288``,
``303-307``), with the ``rerank`` flag plumbed through ``validateModels.setParameters``
(``validateModels.py:28-31``).  ``torchreid`` is not vendored, so the published
algorithm of ``torchreid/utils/rerank.py::re_ranking`` -- Zhong et al., "Re-ranking Person
Re-identification with k-reciprocal Encoding", CVPR 2017, in the widely copied numpy form --
is restated here statement by statement.  **PARITY UNPINNED**: the reference holds no test or
fixture for it and the module cannot be executed here.  The restatement is pinned by its own
invariants (``tests/test_oracle.py``): rows of V sum to 1, the Jaccard distance of a sample with
itself is 0, lambda = 1 returns the normalised squared input, the result is invariant to a
gallery permutation on tie-free inputs.

Canonical tie order: upstream uses ``np.argsort`` (unstable default); as everywhere in this
repo the canonical order is the stable one (value ascending, index ascending).
"""
from __future__ import annotations

import numpy as np

__all__ = ["re_ranking", "re_ranking_details"]


def re_ranking_details(q_g_dist, q_q_dist, g_g_dist, k1=20, k2=6, lambda_value=0.3):
    q_g_dist = np.asarray(q_g_dist, dtype=np.float32)
    q_q_dist = np.asarray(q_q_dist, dtype=np.float32)
    g_g_dist = np.asarray(g_g_dist, dtype=np.float32)
    # "The following naming, e.g. gallery_num, is different from outer scope."
    original_dist = np.concatenate(
        [np.concatenate([q_q_dist, q_g_dist], axis=1),
         np.concatenate([q_g_dist.T, g_g_dist], axis=1)], axis=0)
    original_dist = np.power(original_dist, 2).astype(np.float32)
    original_dist = np.transpose(1. * original_dist / np.max(original_dist, axis=0))
    V = np.zeros_like(original_dist).astype(np.float32)
    initial_rank = np.argsort(original_dist, kind="stable").astype(np.int32)

    query_num = q_g_dist.shape[0]
    gallery_num = q_g_dist.shape[0] + q_g_dist.shape[1]
    all_num = gallery_num
    half = int(np.around(k1 / 2.))

    for i in range(all_num):
        # k-reciprocal neighbors
        forward_k_neigh_index = initial_rank[i, :k1 + 1]
        backward_k_neigh_index = initial_rank[forward_k_neigh_index, :k1 + 1]
        fi = np.where(backward_k_neigh_index == i)[0]
        k_reciprocal_index = forward_k_neigh_index[fi]
        k_reciprocal_expansion_index = k_reciprocal_index
        for j in range(len(k_reciprocal_index)):
            candidate = k_reciprocal_index[j]
            candidate_forward_k_neigh_index = initial_rank[candidate, :half + 1]
            candidate_backward_k_neigh_index = initial_rank[candidate_forward_k_neigh_index, :half + 1]
            fi_candidate = np.where(candidate_backward_k_neigh_index == candidate)[0]
            candidate_k_reciprocal_index = candidate_forward_k_neigh_index[fi_candidate]
            if len(np.intersect1d(candidate_k_reciprocal_index, k_reciprocal_index)) > \
                    2. / 3 * len(candidate_k_reciprocal_index):
                k_reciprocal_expansion_index = np.append(k_reciprocal_expansion_index,
                                                         candidate_k_reciprocal_index)

        k_reciprocal_expansion_index = np.unique(k_reciprocal_expansion_index)
        weight = np.exp(-original_dist[i, k_reciprocal_expansion_index])
        V[i, k_reciprocal_expansion_index] = 1. * weight / np.sum(weight)
    original_dist = original_dist[:query_num, ]
    V0 = V
    if k2 != 1:
        V_qe = np.zeros_like(V, dtype=np.float32)
        for i in range(all_num):
            V_qe[i, :] = np.mean(V[initial_rank[i, :k2], :], axis=0)
        V = V_qe
        del V_qe
    invIndex = []
    for i in range(gallery_num):
        invIndex.append(np.where(V[:, i] != 0)[0])

    jaccard_dist = np.zeros_like(original_dist, dtype=np.float32)

    for i in range(query_num):
        temp_min = np.zeros(shape=[1, gallery_num], dtype=np.float32)
        indNonZero = np.where(V[i, :] != 0)[0]
        indImages = [invIndex[ind] for ind in indNonZero]
        for j in range(len(indNonZero)):
            temp_min[0, indImages[j]] = temp_min[0, indImages[j]] + \
                np.minimum(V[i, indNonZero[j]], V[indImages[j], indNonZero[j]])
        jaccard_dist[i] = 1 - temp_min / (2. - temp_min)

    final_dist = jaccard_dist * (1 - lambda_value) + original_dist * lambda_value
    final_dist = final_dist[:query_num, query_num:]
    return dict(final=final_dist, original=original_dist, initial_rank=initial_rank, V0=V0, V=V,
                jaccard=jaccard_dist)


def re_ranking(q_g_dist, q_q_dist, g_g_dist, k1=20, k2=6, lambda_value=0.3):
    """``torchreid.utils.re_ranking(distmat, distmat_qq, distmat_gg)`` -> ``[Q, G]`` float32."""
    return re_ranking_details(q_g_dist, q_q_dist, g_g_dist, k1, k2, lambda_value)["final"]

"""daliid_b200 -- B200-native retrieval-evaluation hot path of DaliID.

Query x gallery distance matrix, per-query ranking with same-identity / same-camera junk
masking, CMC/mAP, closed-set top-k identification and multi-model distance fusion, behind
the reference's own Python entry points and a C-ABI library of hand-written sm_100a
kernels (``include/daliid_b200.h``).  See DESIGN.md and INTEGRATION.md.
"""
from .metrics import (canonicalize_labels, compute_distance_matrix, evaluate_features,
                      evaluate_rank, evaluate_rank_detailed, fuse_distmats, normalize,
                      re_ranking, topk_features, topk_identify)

__all__ = [
    "canonicalize_labels", "compute_distance_matrix", "evaluate_features", "evaluate_rank",
    "evaluate_rank_detailed", "fuse_distmats", "normalize", "re_ranking", "topk_features",
    "topk_identify",
]
__version__ = "0.1.0"

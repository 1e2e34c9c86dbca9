"""Golden vectors of k-reciprocal re-ranking: inputs and the output of the restated
numpy form (oracle/rerank_oracle.py) -- torchreid itself is not vendored, SURVEY 8c.

    python tests/golden/make_golden_rerank.py     # in the build container; writes rerank.npz
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import rerank_oracle as rr  # noqa: E402


def main():
    g = torch.Generator().manual_seed(12)
    Q, G, D = 23, 90, 16
    cent = torch.randn(9, D, generator=g)
    q = cent[torch.randint(0, 9, (Q,), generator=g)] + 0.7 * torch.randn(Q, D, generator=g)
    x = cent[torch.randint(0, 9, (G,), generator=g)] + 0.7 * torch.randn(G, D, generator=g)
    x[11] = x[4]  # exact duplicates: ties in the neighbour lists
    q = q / torch.norm(q, dim=1, keepdim=True)
    x = x / torch.norm(x, dim=1, keepdim=True)

    def sq(a, b):  # torchreid's compute_distance_matrix(.., "euclidean"): squared, addmm form
        return ((a * a).sum(1)[:, None] + (b * b).sum(1)[None, :] - 2.0 * a @ b.T).numpy()

    qg = (1.0 - torch.mm(q, x.T)).numpy()
    qq, gg = sq(q, q), sq(x, x)
    out = dict(qg=qg, qq=qq, gg=gg)
    for name, (k1, k2, lam) in dict(a=(8, 3, 0.3), b=(20, 6, 0.3), c=(5, 1, 0.7)).items():
        out["final_" + name] = rr.re_ranking(qg, qq, gg, k1, k2, lam)
        out["params_" + name] = np.array([k1, k2, lam], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "rerank.npz"), **out)
    print("rerank.npz", os.path.getsize(os.path.join(HERE, "rerank.npz")))


if __name__ == "__main__":
    main()

"""One process, two GPUs (not a pytest test: needs a 2-GPU box).
    gpurun --gpus 2 -- python tests/run_two_devices.py
Every context sets its kernels' opt-in shared-memory limits on its own device; results on cuda:1
must equal those on cuda:0 bit for bit."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from daliid_b200 import metrics, synth  # noqa: E402


def main():
    assert torch.cuda.device_count() >= 2, "needs two GPUs"
    qf, gf, qp, gp, qc, gc = synth.make_config("market_vit", device="cpu")
    res = []
    for dev in (0, 1, 0):
        q, g = qf.cuda(dev), gf.cuda(dev)
        with torch.cuda.device(dev):
            cmc, mAP = metrics.evaluate_features(q, g, qp, gp, qc, gc)
            d = metrics.compute_distance_matrix(q, g, "cosine")
            vals, ids = metrics.topk_features(q, g, k=20)
            rr = metrics.re_ranking(d[:200, :900], metrics.compute_distance_matrix(q[:200], q[:200], "sqeuclidean", normalize=True),
                                    metrics.compute_distance_matrix(g[:900], g[:900], "sqeuclidean", normalize=True))
            sim = (1.0 - d).contiguous()
            fused = metrics.mrfuse([sim, sim], 20)
        res.append((np.asarray(cmc), mAP, d.cpu(), vals.cpu(), ids.cpu(), rr.cpu(), fused.cpu()))
    for r in res[1:]:
        assert np.array_equal(r[0], res[0][0]) and r[1] == res[0][1]
        for a, b in zip(r[2:], res[0][2:]):
            assert torch.equal(a, b)
    print(f"two devices in one process: identical results (mAP {res[0][1]:.6f})")


if __name__ == "__main__":
    main()

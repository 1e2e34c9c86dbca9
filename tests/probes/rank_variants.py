"""Times the rank stage alone (device-resident distance matrix) with the library's kernel timers.
usage: [DALIID_B200_LIB=...] python tests/probes/rank_variants.py [config ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from daliid_b200 import _lib, metrics, synth
ctx = _lib.get_ctx(0)
tag = os.environ.get("TAG", os.path.basename(os.environ.get("DALIID_B200_LIB", "default")))
for name in (sys.argv[1:] or ["market_resnet50", "market_vit"]):
    over = {"n_ids": int(os.environ["PROBE_NIDS"])} if os.environ.get("PROBE_NIDS") else {}
    qf, gf, qp, gp, qc, gc = synth.make_config(name, device="cuda", **over)
    d = metrics.compute_distance_matrix(qf, gf, metric="cosine")
    for _ in range(3):
        cmc, mAP = metrics.evaluate_rank(d, qp, gp, qc, gc, max_rank=50)[:2]
    torch.cuda.synchronize()
    ctx.timing_enable(True); ctx.timing_reset()
    n = 30
    for _ in range(n):
        cmc, mAP = metrics.evaluate_rank(d, qp, gp, qc, gc, max_rank=50)[:2]
    kt = ctx.timing_read(); ctx.timing_enable(False)
    ms = {k: round(v[1] / n, 4) for k, v in kt.items() if v[0]}
    gb = d.shape[0] * d.shape[1] * 4 / 1e9
    print(f"{tag:28s} {name:16s} {ms}  count {gb / (ms['rank_count'] * 1e-3):7.0f} GB/s  mAP {mAP:.6f}", flush=True)
    del d

"""Three-kernel rank stage (gather / count / finalize: the sharded building blocks) on one GPU (diagnostic)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from daliid_b200 import _lib, metrics, synth, sharded
ctx = _lib.get_ctx(0)
for name in (sys.argv[1:] or ["deepchange", "market_vit"]):
    qf, gf, qp, gp, qc, gc = synth.make_config(name, device="cuda")
    d = metrics.compute_distance_matrix(qf, gf, "cosine")
    ops = sharded.CudaOps(0)
    def run():
        return sharded.evaluate_rank_sharded(d, 0, qp, gp, qc, gc, ops=ops)
    for _ in range(2): run()
    torch.cuda.synchronize()
    ctx.timing_enable(True); ctx.timing_reset()
    for _ in range(5): cmc, mAP = run()[:2]
    kt = ctx.timing_read(); ctx.timing_enable(False)
    print(name, {k: round(v[1] / 5, 4) for k, v in kt.items() if v[0]}, "mAP", mAP, flush=True)

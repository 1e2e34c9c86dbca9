"""Small-sized run of every kernel family for compute-sanitizer (diagnostic; one tool per call):
contraction in every arithmetic mode (store / filter / band / count epilogues), operand preparation,
rank counting (16-bit and byte counters, fused single launch and the three-kernel form), finalize,
top-k and its compaction, fusion, re-ranking, meta-recognition fusion, whole-row ordering."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from daliid_b200 import _lib, metrics, sharded, synth  # noqa: E402

qf, gf, qp, gp, qc, gc = synth.make_features(300, 2100, 200, 120, 5, 2.0, seed=12, device="cuda")
ctx = _lib.get_ctx(0)
for prec in ("fp32", "tf32x3", "tf32c", "tf32", "f16x3", "f16"):
    cmc, mAP = metrics.evaluate_features(qf, gf, qp, gp, qc, gc, precision=prec)
    print(prec, "matrix path", mAP, flush=True)
ctx.fused_count_enable(True)
for prec in ("tf32c", "tf32", "f16x3", "f16"):
    cmc, mAP = metrics.evaluate_features(qf, gf, qp, gp, qc, gc, precision=prec)
    print(prec, "fused counting", mAP, ctx.fused_count_calls(), flush=True)
ctx.fused_count_enable(False)
# many positives per query: byte counters (single launch) and the split three-kernel form
q2, g2, qp2, gp2, qc2, gc2 = synth.make_features(64, 3000, 64, 12, 3, 2.0, seed=3, device="cuda")
print("byte counters", metrics.evaluate_features(q2, g2, qp2, gp2, qc2, gc2)[1], flush=True)
d = metrics.compute_distance_matrix(qf, gf, "cosine")
ops = sharded.CudaOps()
plan = ops.plan(qp, gp, qc, gc)
keys = ops.gather_keys(plan, d[:, :1000].contiguous(), 0) + ops.gather_keys(plan, d[:, 1000:].contiguous(), 1000)
counts = ops.count(plan, d[:, :1000].contiguous(), 0, keys) + ops.count(plan, d[:, 1000:].contiguous(), 1000, keys)
print("slabs", ops.finalize(plan, keys, counts, 300, 2100, 50, "cy_f32")[1], flush=True)
ops.plan_destroy(plan)
v, i = metrics.topk_identify(d, k=20)
fv, fi = metrics.topk_features(qf, gf, k=20)
assert torch.equal(fi, i)
metrics.topk_features(torch.randn(700, 64, device="cuda"), torch.randn(9000, 64, device="cuda"), k=5)  # multi-chunk filter
print("fuse", float(metrics.fuse_distmats([d, d.clone(), d.clone()]).sum()), flush=True)
qq = metrics.compute_distance_matrix(qf, qf, "sqeuclidean", normalize=True)
gg = metrics.compute_distance_matrix(gf, gf, "sqeuclidean", normalize=True)
print("rerank", float(metrics.re_ranking(d, qq, gg).sum()), flush=True)
sim = (1.0 - d).contiguous()
print("mrfuse", float(metrics.mrfuse([sim, sim.clone()], 20).sum()), flush=True)
print("argsort", int(metrics.argsort_rows(d[:50]).sum()), flush=True)
torch.cuda.synchronize()
print("done")

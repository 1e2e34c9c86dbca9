"""Top-k from a materialised matrix (a7): the select kernel against the chunked streaming path (diagnostic)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from daliid_b200 import _lib, metrics, synth
ctx = _lib.get_ctx(0)
qf, gf, *_ = synth.make_config("market_vit", device="cuda")
d = metrics.compute_distance_matrix(qf, gf, "cosine")
for k in (20, 5, 128):
    for _ in range(3):
        v, i = metrics.topk_identify(d, k=k)
    torch.cuda.synchronize()
    ctx.timing_enable(True); ctx.timing_reset()
    for _ in range(10):
        v, i = metrics.topk_identify(d, k=k)
    kt = ctx.timing_read(); ctx.timing_enable(False)
    ref = torch.argsort(d, dim=1, stable=True)[:, :k]
    print(f"{os.environ.get('TAG','')} k={k}: topk kernels {kt['topk'][1] / 10:.4f} ms ({kt['topk'][0] // 10} launches), equal to the stable argsort prefix: {bool(torch.equal(i.long(), ref))}, fallbacks {ctx.fallback_count()}", flush=True)

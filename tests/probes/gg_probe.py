import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from daliid_b200 import _lib, metrics, synth
qf, gf, *_ = synth.make_config("market_vit", device="cuda")
ctx = _lib.get_ctx(0)
def t(fn, n=5):
    fn(); fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for metric in ("cosine", "sqeuclidean"):
    for pad in (True, False):
        ctx.timing_enable(True); ctx.timing_reset()
        ms = t(lambda: metrics.compute_distance_matrix(gf, gf, metric, normalize=True, padded=pad))
        kt = {k: round(v[1] / max(v[0], 1), 3) for k, v in ctx.timing_read().items() if v[0]}
        ctx.timing_enable(False)
        print(metric, "padded" if pad else "contiguous", f"{ms:.3f} ms", kt, flush=True)

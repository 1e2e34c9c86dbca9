"""C5 slab timing probe: fused distance + top-20, 100k x 125k x 512 (diagnostic)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from daliid_b200 import _lib, metrics
Q, G, D, k = 100000, 125000, 512, 20
g = torch.Generator(device="cuda").manual_seed(12)
qf = torch.randn(Q, D, generator=g, device="cuda")
gf = torch.randn(G, D, generator=g, device="cuda")
ctx = _lib.get_ctx(0)
for mode in ("f16x3", "f16"):
    for _ in range(2):
        v, i = metrics.topk_features(qf, gf, k=k, precision=mode)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0 = ctx.fallback_count()
    e0.record()
    for _ in range(3):
        v, i = metrics.topk_features(qf, gf, k=k, precision=mode)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    ctx.timing_enable(True); ctx.timing_reset()
    metrics.topk_features(qf, gf, k=k, precision=mode)
    kt = {kk: (vv[0], round(vv[1], 3)) for kk, vv in ctx.timing_read().items() if vv[0]}
    ctx.timing_enable(False)
    tf = 2.0 * Q * G * D / (ms * 1e-3) / 1e12
    print(f"{mode}: {ms:.2f} ms/eval {tf:.1f} TFLOP/s ({tf / 1623.1:.4f} of bf16 peak) fallbacks {ctx.fallback_count() - f0} kernels {kt}", flush=True)

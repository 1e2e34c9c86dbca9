"""Where a device-resident evaluation step spends its time beyond the kernels (diagnostic).
usage: python tests/probes/step_overhead.py [config]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from daliid_b200 import _lib, metrics, synth
name = sys.argv[1] if len(sys.argv) > 1 else "market_resnet50"
qf, gf, qp, gp, qc, gc = synth.make_config(name, device="cuda")
ctx = _lib.get_ctx(0)
def step():
    return metrics.evaluate_features(qf, gf, qp, gp, qc, gc, precision="f16x3")
for _ in range(5):
    step()
torch.cuda.synchronize()
n = 50
t0 = time.perf_counter()
for _ in range(n):
    step()
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / n * 1e3
ctx.timing_enable(True); ctx.timing_reset()
for _ in range(8):
    step()
kt = ctx.timing_read(); ctx.timing_enable(False)
ks = {k: round(v[1] / 8, 4) for k, v in kt.items() if v[0]}
print(f"{name}: wall {wall:.4f} ms/step, kernels {ks} sum {sum(ks.values()):.4f} ms, beyond kernels {wall - sum(ks.values()):.4f} ms")
# python-side share: the same call with the C entry point stubbed out is not possible; time the label canonicalisation
t0 = time.perf_counter()
for _ in range(200):
    metrics.canonicalize_labels(qp, gp); metrics.canonicalize_labels(qc, gc)
print(f"label canonicalisation {(time.perf_counter() - t0) / 200 * 1e3:.4f} ms/step")

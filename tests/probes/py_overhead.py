import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from daliid_b200 import _lib, metrics, synth
qf, gf, qp, gp, qc, gc = synth.make_config("market_resnet50", device="cuda")
ctx = _lib.get_ctx(0)
real = ctx.lib.dali_eval_features_f32
acc = {"c": 0.0, "n": 0}
def timed(*a):
    t0 = time.perf_counter(); r = real(*a); acc["c"] += time.perf_counter() - t0; acc["n"] += 1; return r
class L:  # proxy
    def __getattr__(self, k): return timed if k == "dali_eval_features_f32" else getattr(ctx_lib, k)
ctx_lib = ctx.lib
ctx.lib = L()
for _ in range(5): metrics.evaluate_features(qf, gf, qp, gp, qc, gc)
torch.cuda.synchronize(); acc["c"] = 0.0; acc["n"] = 0
t0 = time.perf_counter()
for _ in range(50): metrics.evaluate_features(qf, gf, qp, gp, qc, gc)
tot = time.perf_counter() - t0
print(f"per call: total {tot/50*1e3:.3f} ms, inside C {acc['c']/50*1e3:.3f} ms, python {(tot-acc['c'])/50*1e3:.3f} ms")

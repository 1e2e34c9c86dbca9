import os, sys, time
os.environ["DALI_TRACE"] = "1"
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from daliid_b200 import _lib, metrics, synth
what = sys.argv[1] if len(sys.argv) > 1 else "market_vit"
qf, gf, qp, gp, qc, gc = synth.make_config(what, device="cuda")
ctx = _lib.get_ctx(0)
for _ in range(5):
    metrics.evaluate_features(qf, gf, qp, gp, qc, gc)
torch.cuda.synchronize()
ctx.timing_enable(True); ctx.timing_reset()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
metrics.evaluate_features(qf, gf, qp, gp, qc, gc)
e1.record(); t1 = time.perf_counter()
torch.cuda.synchronize()
print(f"host wall {1e6 * (t1 - t0):.1f} us, device events around the call {1e3 * e0.elapsed_time(e1):.1f} us", file=sys.stderr)
ctx.timing_read()
ctx.timing_enable(False)
for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    metrics.evaluate_features(qf, gf, qp, gp, qc, gc)
    t1 = time.perf_counter()
    print(f"python wall of one untimed call {1e6 * (t1 - t0):.1f} us", file=sys.stderr)
# untimed steady state
ctx.timing_enable(False)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(200):
    metrics.evaluate_features(qf, gf, qp, gp, qc, gc)
torch.cuda.synchronize(); t1 = time.perf_counter()
print(f"steady state {1e6 * (t1 - t0) / 200:.1f} us per evaluation", file=sys.stderr)

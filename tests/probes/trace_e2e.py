import os, sys, time
os.environ["DALI_TRACE"] = "1"
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from daliid_b200 import _lib, metrics, synth
what = sys.argv[1] if len(sys.argv) > 1 else "market_resnet50"
qf, gf, qp, gp, qc, gc = synth.make_config(what, device="cpu")
qf, gf = qf.pin_memory(), gf.pin_memory()
ctx = _lib.get_ctx(0)
dq = torch.empty(qf.shape, device="cuda"); dg = torch.empty(gf.shape, device="cuda")
for _ in range(3):
    metrics.evaluate_features(qf, gf, qp, gp, qc, gc)
torch.cuda.synchronize()
# raw copy of the same bytes
for _ in range(2):
    t0 = time.perf_counter(); dq.copy_(qf, non_blocking=True); dg.copy_(gf, non_blocking=True); torch.cuda.synchronize()
    print(f"raw pinned H2D {1e6 * (time.perf_counter() - t0):.0f} us", file=sys.stderr)
ctx.timing_enable(True); ctx.timing_reset()
t0 = time.perf_counter()
metrics.evaluate_features(qf, gf, qp, gp, qc, gc)
t1 = time.perf_counter()
print(f"host wall (timers on) {1e6 * (t1 - t0):.0f} us", file=sys.stderr)
ctx.timing_read(); ctx.timing_enable(False)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20):
    metrics.evaluate_features(qf, gf, qp, gp, qc, gc)
t1 = time.perf_counter()
print(f"steady state e2e {1e6 * (t1 - t0) / 20:.0f} us per evaluation", file=sys.stderr)

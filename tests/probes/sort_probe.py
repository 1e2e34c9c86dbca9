"""Time of the whole-row ordering (dali_argsort_f32) at the Market matrix and one long row (diagnostic)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from daliid_b200 import metrics
for Q, G in [(3368, 15913), (1, 1 << 20), (1, 12936)]:
    d = torch.randn(Q, G, device="cuda")
    for _ in range(2):
        o = metrics.argsort_rows(d)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        o = metrics.argsort_rows(d)
    e1.record(); torch.cuda.synchronize()
    ours = e0.elapsed_time(e1) / 5
    e0.record()
    for _ in range(5):
        r = torch.argsort(d, dim=1, stable=True)
    e1.record(); torch.cuda.synchronize()
    print(f"{Q} x {G}: dali_argsort_f32 {ours:.3f} ms, torch.argsort(stable) {e0.elapsed_time(e1) / 5:.3f} ms, equal {bool(torch.equal(o.long(), r))}")

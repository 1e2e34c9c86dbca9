"""Where the e2e step (pinned host features in, CMC/mAP out) spends its time: raw PCIe copy of the
same bytes vs the library call."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from daliid_b200 import _lib, metrics, synth
qf, gf, qp, gp, qc, gc = synth.make_config("market_resnet50", device="cpu")
qh, gh = qf.pin_memory(), gf.pin_memory()
qd, gd = torch.empty_like(qf, device="cuda"), torch.empty_like(gf, device="cuda")
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
def raw():
    qd.copy_(qh, non_blocking=True); gd.copy_(gh, non_blocking=True)
print(f"raw H2D of q+g ({(qf.numel()+gf.numel())*4/1e6:.1f} MB): {timeit(raw):.3f} ms")
print(f"e2e call (host pinned): {timeit(lambda: metrics.evaluate_features(qh, gh, qp, gp, qc, gc)):.3f} ms")
print(f"device-resident call: {timeit(lambda: metrics.evaluate_features(qd, gd, qp, gp, qc, gc)):.3f} ms")
def split():
    qd.copy_(qh, non_blocking=True); gd.copy_(gh, non_blocking=True)
    return metrics.evaluate_features(qd, gd, qp, gp, qc, gc)
print(f"copy then device call (no overlap): {timeit(split):.3f} ms")

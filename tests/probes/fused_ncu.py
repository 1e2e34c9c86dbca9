"""Minimal driver for ncu: a few fused evaluations of one config (diagnostic)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from daliid_b200 import _lib, metrics, synth
name = sys.argv[1] if len(sys.argv) > 1 else "market_resnet50"
fused = (sys.argv[2] != "0") if len(sys.argv) > 2 else True
qf, gf, qp, gp, qc, gc = synth.make_config(name, device="cuda")
ctx = _lib.get_ctx(0)
ctx.fused_count_enable(fused)
for _ in range(3):
    cmc, mAP = metrics.evaluate_features(qf, gf, qp, gp, qc, gc, precision="f16x3")
torch.cuda.synchronize()
print(name, "fused" if fused else "matrix", mAP)

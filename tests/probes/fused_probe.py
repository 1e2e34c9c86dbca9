"""Timing probe of the fused distance + counting path against the matrix path (diagnostic)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from daliid_b200 import _lib, metrics, synth

ctx = _lib.get_ctx(0)
for name in sys.argv[1:] or ("market_vit", "market_resnet50"):
    qf, gf, qp, gp, qc, gc = synth.make_config(name, device="cuda")
    Q, G = qf.shape[0], gf.shape[0]
    for fused in (True, False):
        ctx.fused_count_enable(fused)
        for _ in range(3):
            metrics.evaluate_features(qf, gf, qp, gp, qc, gc, precision="f16x3")
        torch.cuda.synchronize()
        n0 = ctx.fused_count_calls()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            cmc, mAP = metrics.evaluate_features(qf, gf, qp, gp, qc, gc, precision="f16x3")
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        ctx.timing_enable(True); ctx.timing_reset()
        for _ in range(5):
            metrics.evaluate_features(qf, gf, qp, gp, qc, gc, precision="f16x3")
        kt = {k: (v[0] // 5, round(v[1] / 5, 4)) for k, v in ctx.timing_read().items() if v[0]}
        ctx.timing_enable(False)
        print(f"{name} fused={fused} (took fused path {ctx.fused_count_calls() - n0}/20): {ms:.4f} ms/eval "
              f"{Q * G / ms / 1e6:.1f} Gpairs/s mAP={mAP:.6f} kernels/eval (launches, ms): {kt}", flush=True)
    ctx.fused_count_enable(False)

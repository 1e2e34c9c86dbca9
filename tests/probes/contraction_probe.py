"""Contraction time alone (compute_distance_matrix on device-resident features), per config (diagnostic).
usage: [DALIID_B200_LIB=...] python tests/probes/contraction_probe.py [config ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from daliid_b200 import _lib, metrics, synth
ctx = _lib.get_ctx(0)
tag = os.environ.get("TAG", os.path.basename(os.environ.get("DALIID_B200_LIB", "default")))
for name in (sys.argv[1:] or ["market_vit", "market_resnet50"]):
    qf, gf, *_ = synth.make_config(name, device="cuda")
    for _ in range(3):
        d = metrics.compute_distance_matrix(qf, gf, "cosine", "f16x3")
    torch.cuda.synchronize()
    ctx.timing_enable(True); ctx.timing_reset()
    n = 20
    for _ in range(n):
        d = metrics.compute_distance_matrix(qf, gf, "cosine", "f16x3")
    kt = ctx.timing_read(); ctx.timing_enable(False)
    ms = kt["distmat"][1] / n
    Q, D = qf.shape; G = gf.shape[0]
    print(f"{tag:24s} {name:16s} contraction {ms:.4f} ms  {2.0 * Q * G * D / (ms * 1e-3) / 1e12:6.1f} TFLOP/s algorithmic", flush=True)

import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from daliid_b200 import metrics, synth
qf, gf, *_ = synth.make_config("market_resnet50", device="cuda")
for _ in range(3):
    d = metrics.compute_distance_matrix(qf, gf, "cosine", "fp32", padded=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    d = metrics.compute_distance_matrix(qf, gf, "cosine", "fp32", padded=True)
e1.record(); torch.cuda.synchronize()
print("fp32 distmat ms", e0.elapsed_time(e1) / 5)

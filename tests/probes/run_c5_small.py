import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from daliid_b200 import metrics
g = torch.Generator(device="cuda").manual_seed(12)
qf = torch.randn(16384, 512, generator=g, device="cuda")
gf = torch.randn(262144, 512, generator=g, device="cuda")
for _ in range(3):
    v, i = metrics.topk_features(qf, gf, k=20)
torch.cuda.synchronize()
print("ok", int(i.max()))

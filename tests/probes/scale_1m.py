"""Scale sanity: 1M-row gallery on one GPU (BASELINE config 5's global gallery), fused top-20 for
4096 queries, checked against torch on a few rows."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from daliid_b200 import metrics, _lib
g = torch.Generator(device="cuda").manual_seed(1)
Q, G, D = 4096, 1_000_000, 512
qf = torch.randn(Q, D, generator=g, device="cuda")
gf = torch.randn(G, D, generator=g, device="cuda")
gf[999_999] = qf[7]          # the very last gallery row is query 7 itself
gf[123_456] = qf[7]          # and an exact duplicate earlier: tie broken by index
v, i = metrics.topk_features(qf, gf, k=20)
torch.cuda.synchronize()
assert int(i[7, 0]) == 123_456 and int(i[7, 1]) == 999_999, i[7, :3]
qn = qf[:16] / qf[:16].norm(dim=1, keepdim=True)
gn = gf / gf.norm(dim=1, keepdim=True)
ref = 1.0 - qn @ gn.T
rv, ri = torch.topk(ref, 20, dim=1, largest=False)
same = (ri == i[:16].long()).float().mean().item()
print("fallbacks", _lib.get_ctx(0).fallback_count(), "agreement with torch fp32 top-20 on 16 rows:", same,
      "max |dv|", (rv - v[:16]).abs().max().item())
assert same > 0.97 and (rv - v[:16]).abs().max().item() < 2e-5
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); metrics.topk_features(qf, gf, k=20); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"{Q} x {G} x {D}: {ms:.1f} ms, {2*Q*G*D/ms/1e9:.0f} TFLOP/s")

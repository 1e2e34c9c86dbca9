"""Per-step cost of the fused-mean contraction launches (diagnostic)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from daliid_b200 import _lib, metrics, synth
from daliid_b200._lib import c_vp
name = sys.argv[1] if len(sys.argv) > 1 else "market_vit"
qf, gf, *_ = synth.make_config(name, device="cuda")
Q, D = qf.shape; G = gf.shape[0]
ctx = _lib.get_ctx(0); ctx.attach_torch_stream()
lda = (G + 7) // 8 * 8
acc = torch.empty((Q, lda), device="cuda"); out = torch.empty((Q, (G + 3) // 4 * 4), device="cuda")
prec = metrics._precision("f16x3", True, D)
def run(step, n, with_out):
    ctx.check(ctx.lib.dali_distmat_fuse_mean_f32(ctx.h, c_vp(qf.data_ptr()), Q, c_vp(gf.data_ptr()), G, D, 0, prec, 1,
              c_vp(out.data_ptr() if with_out else None), out.shape[1], c_vp(acc.data_ptr()), lda, step, n))
for label, step, n, wo in [("acc = d", 0, 3, False), ("acc += d", 1, 3, False), ("acc = (acc+d)/3", 2, 3, False),
                           ("acc = (acc+d)/2", 1, 2, False), ("acc += d, own matrix too", 1, 3, True)]:
    for _ in range(3): run(step, n, wo)
    torch.cuda.synchronize()
    ctx.timing_enable(True); ctx.timing_reset()
    for _ in range(10): run(step, n, wo)
    kt = ctx.timing_read(); ctx.timing_enable(False)
    print(f"{name} {label:28s} contraction {kt['distmat'][1] / 10:.4f} ms")
metrics.compute_distance_matrix(qf, gf, "cosine", "f16x3")
ctx.timing_enable(True); ctx.timing_reset()
for _ in range(10): metrics.compute_distance_matrix(qf, gf, "cosine", "f16x3")
kt = ctx.timing_read(); ctx.timing_enable(False)
print(f"{name} plain store                  contraction {kt['distmat'][1] / 10:.4f} ms")

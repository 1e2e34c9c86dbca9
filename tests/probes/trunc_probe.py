"""Accumulator-truncation probe (diagnostic, not a test): error of the f16x3 contraction against a
float64 reference for exact duplicates, near-duplicates, adversarial partial-sum trajectories
(energy of the dot product concentrated at the start / the end of K) and random pairs, per D.

    DALI_F16X3_TWO_PASS=0|1 DALI_F16X3_COMP=<kappa|0> python tests/probes/trunc_probe.py [time]

The two environment variables are read once per process (calibration: run with COMP=0 and fit
kappa = mean loss / N; validation: run without them)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402

from daliid_b200 import _lib, metrics  # noqa: E402

tp = os.environ.get("DALI_F16X3_TWO_PASS", "default")
comp = os.environ.get("DALI_F16X3_COMP", "default")
g = torch.Generator().manual_seed(5)
for D in (512, 768, 1024, 1536, 2048, 3072, 3840, 4096):
    a = torch.randn(256, D, generator=g)
    head = a[:32].clone()
    head[:, 32:] *= 0.02            # nearly all of |x|^2 in the first k-block
    tail = a[32:64].clone()
    tail[:, :-32] *= 0.02           # ... in the last k-block
    a = torch.cat([a, head, tail])
    b = torch.cat([a[:128], a[:64] + 0.01 * torch.randn(64, D, generator=g), head, tail,
                   torch.randn(320, D, generator=g)])
    an = a.double() / a.double().norm(dim=1, keepdim=True)
    bn = b.double() / b.double().norm(dim=1, keepdim=True)
    ref = 1.0 - an @ bn.T
    out = metrics.compute_distance_matrix(a.cuda(), b.cuda(), "cosine", "f16x3").cpu().double()
    e = out - ref
    dup = torch.stack([e[i, i] for i in range(128)])
    near = torch.stack([e[i, 128 + i] for i in range(64)])
    hd = torch.stack([e[256 + i, 192 + i] for i in range(32)])
    tl = torch.stack([e[288 + i, 224 + i] for i in range(32)])
    rnd = e[:256, 256:]
    two = (D > 768) if tp == "default" else tp != "0"
    n_mma = D / 16 if two else 3 * D / 16
    print(f"two_pass={tp} comp={comp} D={D:5d} dup mean {dup.mean():+.3e} max|.| {dup.abs().max():.3e} | "
          f"near mean {near.mean():+.3e} max {near.abs().max():.3e} | head-heavy mean {hd.mean():+.3e} max "
          f"{hd.abs().max():.3e} | tail-heavy mean {tl.mean():+.3e} max {tl.abs().max():.3e} | random mean "
          f"{rnd.mean():+.2e} max {rnd.abs().max():.3e} | kappa(dup) = {dup.mean().item() / 2**-24 / n_mma:.4f} "
          f"kappa(head) = {hd.mean().item() / 2**-24 / n_mma:.4f}", flush=True)

if len(sys.argv) > 1 and sys.argv[1] == "time":
    from daliid_b200 import synth
    ctx = _lib.get_ctx(0)
    for name in ("market_vit", "market_resnet50"):
        qf, gf, qp, gp, qc, gc = synth.make_config(name, device="cuda")
        for _ in range(3):
            metrics.evaluate_features(qf, gf, qp, gp, qc, gc, precision="f16x3")
        ctx.timing_enable(True)
        ctx.timing_reset()
        for _ in range(10):
            cmc, mAP = metrics.evaluate_features(qf, gf, qp, gp, qc, gc, precision="f16x3")
        kt = {k: round(v[1] / max(v[0], 1), 4) for k, v in ctx.timing_read().items() if v[0]}
        ctx.timing_enable(False)
        print(f"two_pass={tp} {name}: kernel ms {kt} mAP={mAP:.6f}", flush=True)

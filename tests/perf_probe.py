"""Ad-hoc performance probe (not a test): times the path on the other BASELINE shapes.
    python tests/perf_probe.py [market_vit|deepchange|topk|faceid]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from daliid_b200 import _lib, metrics, synth  # noqa: E402


def timeit(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, out


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "market_vit"
    ctx = _lib.get_ctx(0)
    if what in ("market_vit", "deepchange", "market_resnet50"):
        qf, gf, qp, gp, qc, gc = synth.make_config(what, device="cuda")
        Q, G = qf.shape[0], gf.shape[0]
        for prec in ("f16x3", "tf32c", "tf32", "fp32"):
            if prec == "fp32" and what == "deepchange":
                continue
            ctx.timing_enable(True); ctx.timing_reset()
            ms, (cmc, mAP) = timeit(lambda: metrics.evaluate_features(qf, gf, qp, gp, qc, gc, precision=prec))
            kt = {k: round(v[1] / max(v[0], 1), 4) for k, v in ctx.timing_read().items() if v[0]}
            ctx.timing_enable(False)
            print(f"{what} {prec}: {ms:.3f} ms/eval, {Q * G / ms / 1e6:.2f} Gpairs/s, mAP={mAP:.5f} "
                  f"R1={cmc[0]:.4f}, kernel ms: {kt}", flush=True)
        d = metrics.compute_distance_matrix(qf, gf, "cosine", "tf32c")
        ms, _ = timeit(lambda: metrics.topk_identify(d, k=20))
        print(f"{what} topk k=20 from distmat: {ms:.3f} ms -> {4 * Q * G / ms / 1e6:.0f} GB/s", flush=True)
        ms, _ = timeit(lambda: metrics.evaluate_rank(d, qp, gp, qc, gc))
        print(f"{what} evaluate_rank from device distmat: {ms:.3f} ms", flush=True)
        d2 = d.clone()
        ms, _ = timeit(lambda: metrics.fuse_distmats([d, d2]))
        print(f"{what} fuse 2: {ms:.3f} ms -> {12 * Q * G / ms / 1e6:.0f} GB/s", flush=True)
    elif what == "rerank":
        # SURVEY 8f N1 at the Market shape: qq / gg / qg from the contraction, then re-ranking
        qf, gf, qp, gp, qc, gc = synth.make_config("market_vit", device="cuda")
        Q, G = qf.shape[0], gf.shape[0]
        qg = metrics.compute_distance_matrix(qf, gf, "cosine")
        qq = metrics.compute_distance_matrix(qf, qf, "sqeuclidean", normalize=True)
        gg = metrics.compute_distance_matrix(gf, gf, "sqeuclidean", normalize=True)
        for _ in range(3):
            metrics.re_ranking(qg, qq, gg)
        ctx.timing_enable(True); ctx.timing_reset()
        ms, out = timeit(lambda: metrics.re_ranking(qg, qq, gg), n=10, warm=0)
        kt = {k: (v[0], round(v[1], 3)) for k, v in ctx.timing_read().items() if v[0]}
        ctx.timing_enable(False)
        c0, m0 = metrics.evaluate_rank(qg, qp, gp, qc, gc)
        c1, m1 = metrics.evaluate_rank(out, qp, gp, qc, gc)
        print(f"rerank market_vit N={Q + G}: {ms:.2f} ms per call; kernels (launches, total ms): {kt}; "
              f"mAP {m0:.4f} -> {m1:.4f}, R1 {c0[0]:.4f} -> {c1[0]:.4f}", flush=True)
        ms, _ = timeit(lambda: (metrics.compute_distance_matrix(qf, qf, "sqeuclidean", normalize=True),
                                metrics.compute_distance_matrix(gf, gf, "sqeuclidean", normalize=True)), n=3, warm=3)
        print(f"qq + gg distance matrices: {ms:.2f} ms", flush=True)
    elif what == "peaks":
        # measured library peaks beside MEASURED_PEAKS.json (bf16): cuBLAS TF32 and fp32 matmul 8192^3
        n = 8192
        a = torch.randn(n, n, device="cuda"); b = torch.randn(n, n, device="cuda")
        for name, allow in (("tf32", True), ("fp32", False)):
            torch.backends.cuda.matmul.allow_tf32 = allow
            ms, _ = timeit(lambda: a @ b, n=10, warm=3)
            print(f"cuBLAS {name} matmul {n}^3: {ms:.3f} ms -> {2 * n ** 3 / ms / 1e9:.1f} TFLOP/s", flush=True)
        ah, bh = a.bfloat16(), b.bfloat16()
        ms, _ = timeit(lambda: ah @ bh, n=10, warm=3)
        print(f"cuBLAS bf16 matmul {n}^3: {ms:.3f} ms -> {2 * n ** 3 / ms / 1e9:.1f} TFLOP/s", flush=True)
        ah, bh = a.half(), b.half()
        ms, _ = timeit(lambda: ah @ bh, n=10, warm=3)
        print(f"cuBLAS fp16 matmul {n}^3: {ms:.3f} ms -> {2 * n ** 3 / ms / 1e9:.1f} TFLOP/s", flush=True)
        src = torch.empty(1 << 28, device="cuda"); dst = torch.empty_like(src)
        ms, _ = timeit(lambda: dst.copy_(src), n=10, warm=3)
        print(f"device copy 1 GiB: {ms:.3f} ms -> {2 * src.numel() * 4 / ms / 1e6:.0f} GB/s (read + write)", flush=True)
    elif what == "mrfuse":
        # SURVEY 8f N3 at the Market shape: three models' similarity matrices -> Weibull weights -> fusion
        sims = []
        for seed in (12, 13, 14):
            qf, gf, qp, gp, qc, gc = synth.make_config("market_vit", device="cuda")
            g = torch.Generator(device="cuda").manual_seed(seed)
            qf = qf + 0.5 * torch.randn(qf.shape, device="cuda", generator=g)
            gf = gf + 0.5 * torch.randn(gf.shape, device="cuda", generator=g)
            sims.append(metrics.compute_distance_matrix(qf, gf, "dot", normalize=True, padded=False))
        Q, G = sims[0].shape
        ctx.timing_enable(True); ctx.timing_reset()
        ms, fused = timeit(lambda: metrics.mrfuse(sims, 20), n=3, warm=1)
        kt = {k: (v[0], round(v[1], 3)) for k, v in ctx.timing_read().items() if v[0]}
        ctx.timing_enable(False)
        c0, m0 = metrics.evaluate_rank((1.0 - sims[0]).contiguous(), qp, gp, qc, gc)
        c1, m1 = metrics.evaluate_rank((1.0 - fused).float().contiguous(), qp, gp, qc, gc)
        print(f"mrfuse 3 x [{Q},{G}]: {ms:.2f} ms per call; kernels (launches, total ms over 3 calls): {kt}; "
              f"mAP {m0:.4f} (model 0) -> {m1:.4f} (fused)", flush=True)
        # CPU restatement of the reference on a column sample, scaled to all columns
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        from oracle import mrfuse_oracle
        sub = sims[0].cpu()
        t0 = time.time()
        srt, small = mrfuse_oracle.tail_of_columns(sub)
        t1 = time.time()
        cols = slice(0, 1024)
        mrfuse_oracle.weibull_fit(srt[cols] + 1 - small[cols, None])
        t2 = time.time()
        print(f"CPU (oracle, {torch.get_num_threads()} threads): tail selection {t1 - t0:.2f} s per model, "
              f"Weibull fit {t2 - t1:.2f} s per 1024 columns -> ~{(t1 - t0 + (t2 - t1) * G / 1024) * 3:.0f} s "
              f"for three models (without the CDF pass)", flush=True)
    elif what == "ensemble":
        # BASELINE config 4: three models' Market-shaped distance matrices, mean fusion, ranking
        feats = []
        for seed in (12, 13, 14):
            q, g, qp, gp, qc, gc = synth.make_config("market_vit", device="cuda", seed=seed) \
                if "seed" in synth.make_config.__code__.co_varnames else synth.make_config("market_vit", device="cuda")
            feats.append((q + 0.01 * seed, g + 0.01 * seed))
        Q, G = feats[0][0].shape[0], feats[0][1].shape[0]

        def run():
            ds = [metrics.compute_distance_matrix(q, g, "cosine") for q, g in feats]
            fused = metrics.fuse_distmats(ds)
            return metrics.evaluate_rank(fused, qp, gp, qc, gc)
        ctx.timing_enable(True); ctx.timing_reset()
        ms, (cmc, mAP) = timeit(run, n=5, warm=2)
        kt = {k: round(v[1] / 5, 4) for k, v in ctx.timing_read().items() if v[0]}
        ctx.timing_enable(False)
        print(f"ensemble of 3 (market_vit shape): {ms:.3f} ms per run (3 distmats + mean fusion + rank), "
              f"{3 * Q * G / ms / 1e6:.1f} Gpairs/s, mAP={mAP:.4f}; kernel ms per run: {kt}", flush=True)
    elif what == "plan":
        # host cost of the rank plan for an 8-slab global gallery (the sharded path builds it on
        # every rank and every step)
        import numpy as np
        from daliid_b200 import sharded
        ops = sharded.CudaOps(0)
        for world in (1, 8):
            G, Q = 15913 * world, 3368
            rng = np.random.default_rng(0)
            gp = rng.integers(0, 751 * world, G).astype(np.int32); gc = rng.integers(0, 6, G).astype(np.int32)
            qp = rng.integers(0, 751, Q).astype(np.int32); qc = rng.integers(0, 6, Q).astype(np.int32)
            for _ in range(3):
                ops.plan_destroy(ops.plan(qp, gp, qc, gc))
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(20):
                pl = ops.plan(qp, gp, qc, gc)
                ops.plan_destroy(pl)
            t1 = time.perf_counter()
            torch.cuda.synchronize()
            print(f"plan create+destroy, G={G}: {(t1 - t0) / 20 * 1e3:.3f} ms host per call", flush=True)
    elif what == "faceid_slab":
        # BASELINE config 5 as one of 8 GPUs sees it: all 100k queries against a 125k gallery slab
        Q, G, D = 100000, 125000, 512
        g = torch.Generator(device="cuda").manual_seed(12)
        qf = torch.randn(Q, D, generator=g, device="cuda")
        gf = torch.randn(G, D, generator=g, device="cuda")
        for prec in sys.argv[2:] or ("f16x3",):
            ctx.timing_enable(True); ctx.timing_reset()
            n_fb = ctx.fallback_count()
            ms, (v, i) = timeit(lambda: metrics.topk_features(qf, gf, k=20, precision=prec), n=1, warm=1)
            kt = {k: (v[0], round(v[1], 3)) for k, v in ctx.timing_read().items() if v[0]}
            ctx.timing_enable(False)
            print(f"faceid_slab {Q}x{G} D={D} {prec}: {ms:.1f} ms, {Q * G / ms / 1e6:.2f} Gpairs/s, "
                  f"{2 * Q * G * D / ms / 1e9:.0f} TFLOP/s, fallbacks {ctx.fallback_count() - n_fb}, "
                  f"kernels (launches, total ms): {kt}", flush=True)
    elif what == "faceid":
        # scaled-down BASELINE config 5 (100k x 1M, D=512): 16k x 262k here, k=20
        Q, G, D = 16384, 262144, 512
        g = torch.Generator(device="cuda").manual_seed(12)
        qf = torch.randn(Q, D, generator=g, device="cuda")
        gf = torch.randn(G, D, generator=g, device="cuda")
        for prec in ("f16x3", "tf32c", "tf32"):
            ctx.timing_enable(True); ctx.timing_reset()
            n_fb = ctx.fallback_count()
            ms, (v, i) = timeit(lambda: metrics.topk_features(qf, gf, k=20, precision=prec), n=2, warm=1)
            kt = {k: (v[0], round(v[1], 3)) for k, v in ctx.timing_read().items() if v[0]}
            ctx.timing_enable(False)
            print(f"faceid {Q}x{G} D={D} {prec}: {ms:.1f} ms, {Q * G / ms / 1e6:.2f} Gpairs/s, "
                  f"{2 * Q * G * D / ms / 1e9:.0f} TFLOP/s, fallbacks {ctx.fallback_count() - n_fb}, "
                  f"kernels (launches, total ms): {kt}", flush=True)


if __name__ == "__main__":
    main()

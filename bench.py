#!/usr/bin/env python
"""Benchmark of the retrieval-evaluation hot path (BASELINE.json metric: Q x G pairs/s for
distmat + rank + CMC/mAP).

    python bench.py --gpus N --steps K --warmup W            # this framework (B200)
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path

A "step" is one full evaluation: L2-normalise -> Q x G x D distance matrix -> per-query ranking
with junk masking -> CMC/mAP on the host.  N=1 workload = BASELINE config[1] (Market-1501
shape, ResNet-50 D=2048 features, synthetic seed 12).  N>1 = the same queries against a gallery
sharded over the ranks (one Market-sized slab per GPU, weak scaling), two tiny all-reduces
per step.  One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "QxG pairs/s (distmat+rank+CMC/mAP)"
UNIT = "pairs/s"
WORKLOAD = "market_resnet50"  # BASELINE.json configs[1]


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], bf16=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained"),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """Samples SM clocks / clock-event (throttle) reasons during the timed region: NVML every 10 ms
    from a helper thread (the timed loops last only ~100 ms), nvidia-smi -lms 100 as the fallback."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None
        self.nvml = None
        self.stop_flag = threading.Event()
        self.sm, self.mx, self.reason_bits = [], None, 0

    def _nvml_handle(self):
        import pynvml
        import torch
        pynvml.nvmlInit()
        try:
            uuid = "GPU-" + str(torch.cuda.get_device_properties(self.gpu).uuid)
            h = pynvml.nvmlDeviceGetHandleByUUID(uuid)
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
        return pynvml, h

    def start(self):
        try:
            self.nvml, self.h = self._nvml_handle()
            self.mx = float(self.nvml.nvmlDeviceGetMaxClockInfo(self.h, self.nvml.NVML_CLOCK_SM))
            self.nvml.nvmlDeviceGetClockInfo(self.h, self.nvml.NVML_CLOCK_SM)
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        reasons = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop_flag.is_set():
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)))
                self.reason_bits |= int(reasons(self.h))
            except Exception:
                pass
            self.stop_flag.wait(0.010)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nvml:
            self.stop_flag.set()
            self.t.join(timeout=1)
            n = self.nvml
            bits = {"hw_slowdown": getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                    "hw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                    "sw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                    "sw_power_cap": getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
            reasons = sorted(k for k, b in bits.items() if self.reason_bits & b)
            return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.mx,
                    "reasons": reasons, "samples": len(self.sm), "source": "nvml, 10 ms period"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvidia-smi -lms 100"}


def make_workload(name, rank, world, device):
    """Queries are shared by every rank (same seed); each rank draws its own gallery slab.

    Weak scaling: the gallery grows by one Market-sized slab per GPU and its identity space
    grows with it (n_ids x world identities, as a larger gallery has more people, not more
    images per person), so a query keeps ~21 same-identity gallery items whatever the number
    of slabs and the per-GPU work is fixed.  Queries carry identities of the first n_ids."""
    from daliid_b200 import synth
    cfg = dict(synth.CONFIGS[name])
    Q, G, D = cfg["Q"], cfg["G"], cfg["D"]
    n_ids = cfg["n_ids"]
    q_pid, _, q_cam, _ = synth.make_labels(Q, G, n_ids, cfg["n_cams"], seed=12)
    gen = torch.Generator().manual_seed(12)
    centers = torch.randn(n_ids, D, generator=gen)
    qf = centers[torch.from_numpy(q_pid).long()] + cfg["sigma"] * torch.randn(Q, D, generator=gen)
    if world > 1:  # identities that only exist in the (larger) gallery
        gen2 = torch.Generator().manual_seed(13)
        centers = torch.cat([centers, torch.randn(n_ids * (world - 1), D, generator=gen2)])
    slabs = []
    for r in range(world):
        g = torch.Generator().manual_seed(1000 + r)
        pid = torch.randint(0, n_ids * world, (G,), generator=g)
        cam = torch.randint(0, cfg["n_cams"], (G,), generator=g)
        slabs.append((pid.numpy().astype(np.int32), cam.numpy().astype(np.int32)))
    g = torch.Generator().manual_seed(2000 + rank)
    gf = centers[torch.from_numpy(slabs[rank][0]).long()] + cfg["sigma"] * torch.randn(G, D, generator=g)
    g_pid_all = np.concatenate([s[0] for s in slabs])
    g_cam_all = np.concatenate([s[1] for s in slabs])
    return dict(cfg=cfg, qf=qf.contiguous(), gf=gf.contiguous(), q_pid=q_pid, q_cam=q_cam,
                g_pid_all=g_pid_all, g_cam_all=g_cam_all, g0=rank * G)


# ----------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation of the path
# ----------------------------------------------------------------------------------------
def cpu_reference_step(qf, gf, q_pid, g_pid, q_cam, g_cam):
    """validateModels.py:41-47 verbatim on CPU torch (all host threads) followed by what an
    installed torchreid runs at validateModels.py:68-69: numpy argsort + the compiled
    per-query loop (oracle/rank_oracle.c, the restated Cython evaluator)."""
    from oracle import c_oracle
    qn = qf / torch.norm(qf, dim=1, keepdim=True)
    gn = gf / torch.norm(gf, dim=1, keepdim=True)
    distmat = 1.0 - torch.mm(qn, gn.T)
    distmat = distmat.numpy()
    return c_oracle.evaluate_rank_c(distmat, q_pid, g_pid, q_cam, g_cam, tie="numpy_default")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import c_oracle
    c_oracle.build()
    # torchrun exports OMP_NUM_THREADS=1; the reference arm uses every host core it can
    try:
        torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    except (AttributeError, RuntimeError):
        torch.set_num_threads(max(1, os.cpu_count() or 1))
    wl = make_workload(WORKLOAD, 0, 1, "cpu")
    cfg = wl["cfg"]
    G = cfg["G"]
    pairs = cfg["Q"] * G
    a = (wl["qf"], wl["gf"], wl["q_pid"], wl["g_pid_all"][:G], wl["q_cam"], wl["g_cam_all"][:G])
    for _ in range(args.warmup):
        cpu_reference_step(*a)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cmc, mAP = cpu_reference_step(*a)
    dt = time.perf_counter() - t0
    v = pairs * args.steps / dt
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic (seed 12, class centres + gaussian noise)",
        "config": {"workload": f"{WORKLOAD}: Q={cfg['Q']} G={G} D={cfg['D']} cosine, CPU host cores",
                   "note": "torch.mm uses all host threads; argsort + evaluator loop are single-threaded "
                           "as upstream"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "one Market-sized slab (Q=3368 x G=15913) per step: CPU torch.mm + "
                                   "np.argsort + compiled loop; at N>1 the arm's gallery is N such slabs"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "mAP": mAP, "host_cpus": os.cpu_count(),
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------
def measure_tf32_peak(dev):
    """cuBLAS TF32 8192^3, best of 5 (TFLOP/s): MEASURED_PEAKS.json holds no TF32 figure, so the
    single-pass TF32 mode is rated against this, measured the way the driver measured bf16."""
    n = 8192
    a = torch.randn(n, n, device=dev)
    b = torch.randn(n, n, device=dev)
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        for _ in range(2):
            a @ b
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            a @ b
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    del a, b
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


def measure_c5(dev, rank, world, dist, tf32_peak):
    """BASELINE config 4 beside the headline: 100k queries x 125k gallery rows PER GPU (1M at 8
    GPUs), D=512, fused distance + top-20 (the Q x G matrix is never written), per-slab top-k
    all-gathered and merged.  Device-resident synthetic features, 3 timed evaluations per
    arithmetic mode: the default fp32-class f16x3 (three fp16 passes) and the two single-pass
    modes the 0.01 pp rule admits for identification (f16, tf32), each against ITS tensor peak."""
    from daliid_b200 import sharded
    Q, Gs, D, k = 100000, 125000, 512, 20
    gq = torch.Generator(device=dev).manual_seed(12)
    qf = torch.randn(Q, D, generator=gq, device=dev)
    gg = torch.Generator(device=dev).manual_seed(1000 + rank)
    gf = torch.randn(Gs, D, generator=gg, device=dev)
    pk = peaks()
    out = {"workload": f"Q={Q} x G={Gs}/GPU (global {Gs * world}) x D={D}, fused distance + top-{k}",
           "scaling": "weak (one 125k slab per GPU)"}
    ref_ids = None
    for mode, passes, peak in (("f16x3", 3, pk["bf16"]), ("f16", 1, pk["bf16"]), ("tf32", 1, tf32_peak)):
        for _ in range(2):
            sharded.topk_features_sharded(qf, gf, rank * Gs, k=k, precision=mode)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 3
        e0.record()
        for _ in range(n):
            v, i = sharded.topk_features_sharded(qf, gf, rank * Gs, k=k, precision=mode)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        tf = 2.0 * Q * Gs * D / (ms * 1e-3) / 1e12
        blk = {"ms_per_eval": ms, "pairs_per_s": Q * Gs * world / (ms * 1e-3), "tflops_per_gpu": tf,
               "tensor_passes": passes, "peak_tflops": peak,
               # algorithmic FLOPs over the peak of the MMA kind used; the ceiling is 1 / passes
               "frac_of_peak": tf / peak, "frac_of_own_ceiling": tf * passes / peak}
        if mode == "f16x3":
            ref_ids = i
            blk["frac_of_bf16_peak"] = tf / pk["bf16"]
        else:  # rank-1 agreement of the single-pass mode with the fp32-class result
            blk["top1_agreement_with_f16x3"] = float((i[:, 0] == ref_ids[:, 0]).float().mean().item())
        out[mode] = blk
    out.update({kk: out["f16x3"][kk] for kk in ("ms_per_eval", "pairs_per_s", "tflops_per_gpu", "frac_of_bf16_peak")})
    out["frac_ceiling"] = 1.0 / 3.0
    del qf, gf
    return out


def measure_c1(dev, timed, ctx):
    """BASELINE config 0, the north-star target shape: Market-1501, ViT D=768, full evaluation
    (target: under 50 ms on one B200)."""
    from daliid_b200 import metrics, synth
    qf, gf, qp, gp, qc, gc = synth.make_config("market_vit", device=dev)
    Q, G, D = qf.shape[0], gf.shape[0], qf.shape[1]

    def step():
        return metrics.evaluate_features(qf, gf, qp, gp, qc, gc, precision="f16x3")
    for _ in range(3):
        step()
    ms, (cmc, mAP) = timed(step, 10)
    ctx.timing_enable(True)
    ctx.timing_reset()
    for _ in range(4):
        step()
    kt = ctx.timing_read()
    ctx.timing_enable(False)
    n_dm, ms_dm = kt["distmat"]
    tf = 2.0 * Q * G * D / (ms_dm / max(n_dm, 1) * 1e-3) / 1e12
    return {"workload": f"market_vit: Q={Q} x G={G} x D={D}, cosine, f16x3", "ms_per_step": ms / 10,
            "pairs_per_s": Q * G / (ms / 10 * 1e-3), "target_ms": 50.0, "under_target": ms / 10 < 50.0,
            "kernel_ms_per_step": {k: v[1] / 4 for k, v in kt.items() if v[0]},
            "contraction_tflops": tf, "contraction_frac_of_bf16_peak": tf / peaks()["bf16"],
            "mAP": mAP, "rank1": float(cmc[0])}


def measure_c3(dev, rank, world, dist, timed, ctx):
    """BASELINE config 2: DeepChange shape (17527 x 62956, D=768), STRONG scaling at the run's N --
    the gallery is split over the ranks, every rank holds all queries."""
    from daliid_b200 import metrics, sharded, synth
    qf, gf, qp, gp, qc, gc = synth.make_config("deepchange", device=dev)
    Q, G, D = qf.shape[0], gf.shape[0], qf.shape[1]
    g0, gs = sharded.slab_bounds(G, world, rank)
    slab = gf[g0:g0 + gs].contiguous()
    del gf

    def step():
        if world == 1:
            return metrics.evaluate_features(qf, slab, qp, gp, qc, gc, precision="f16x3")
        return sharded.evaluate_features_sharded(qf, slab, g0, qp, gp, qc, gc, precision="f16x3")
    for _ in range(3):
        step()
    ms, (cmc, mAP) = timed(step, 5)
    ctx.timing_enable(True)
    ctx.timing_reset()
    for _ in range(2):
        step()
    kt = ctx.timing_read()
    ctx.timing_enable(False)
    return {"workload": f"deepchange: Q={Q} x G={G} x D={D}, cosine, f16x3, gallery split over {world} GPU(s)",
            "scaling": "strong", "ms_per_step": ms / 5, "pairs_per_s": Q * G / (ms / 5 * 1e-3),
            "kernel_ms_per_step_rank0": {k: v[1] / 2 for k, v in kt.items() if v[0]}, "mAP": mAP,
            "rank1": float(cmc[0])}


def measure_c4(dev, timed, ctx):
    """BASELINE config 3: three models' Market-shaped distance matrices (seeds 12, 13, 14) -> mean
    fusion in the reference's operation order -> rank / CMC / mAP (evaluate.py:260-279)."""
    from daliid_b200 import metrics, synth
    sets = [synth.make_config("market_vit", seed=sd, device=dev) for sd in (12, 13, 14)]
    _, _, qp, gp, qc, gc = sets[0]
    Q, G = sets[0][0].shape[0], sets[0][1].shape[0]

    def step_unfused():
        ds = [metrics.compute_distance_matrix(q, g, "cosine", "f16x3") for q, g, *_ in sets]
        return metrics.evaluate_rank(metrics.fuse_distmats(ds), qp, gp, qc, gc)

    def step():  # the mean is formed in the three contractions' epilogues (no fusion pass)
        _, mean = metrics.ensemble_distance_matrices([s_[0] for s_ in sets], [s_[1] for s_ in sets], "cosine",
                                                     "f16x3", individual=False)
        return metrics.evaluate_rank(mean, qp, gp, qc, gc)
    for _ in range(3):
        step()
        step_unfused()
    ms_unfused, (cmc_u, mAP_u) = timed(step_unfused, 10)
    ms, (cmc, mAP) = timed(step, 10)
    assert mAP == mAP_u and np.array_equal(cmc, cmc_u), "fused ensemble differs from the separate fusion pass"
    ctx.timing_enable(True)
    ctx.timing_reset()
    for _ in range(4):
        step()
    kt = ctx.timing_read()
    ctx.timing_enable(False)
    return {"workload": f"3 x market_vit (Q={Q} x G={G} x D=768) -> mean fusion -> rank", "ms_per_step": ms / 10,
            "pairs_per_s": Q * G / (ms / 10 * 1e-3), "kernel_ms_per_step": {k: v[1] / 4 for k, v in kt.items() if v[0]},
            "ms_per_step_separate_fusion_pass": ms_unfused / 10,
            "mAP": mAP, "rank1": float(cmc[0])}


def measure_h2d_ceiling(dev, nbytes, world, dist):
    """Raw pinned host -> device copy of one step's input bytes, all ranks at the same time: the
    ceiling of the end-to-end number on this host (aggregate GB/s over the ranks)."""
    h = torch.empty(nbytes // 4, dtype=torch.float32).pin_memory()
    d = torch.empty(nbytes // 4, dtype=torch.float32, device=dev)
    best = 1e9
    for it in range(4):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        d.copy_(h, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        if it:
            best = min(best, ms)
    del h, d
    return {"bytes_per_rank": int(nbytes), "ms": best, "gbs_per_rank": nbytes / (best * 1e-3) / 1e9,
            "gbs_aggregate": world * nbytes / (best * 1e-3) / 1e9}


def stock_gpu_baseline(qf, gf, q_pid, g_pid, q_cam, g_cam, timed):
    """SURVEY 8d's second baseline: the obvious stock-library GPU formulation of the same
    evaluation -- cuBLAS fp32 torch.mm + torch.argsort + vectorised torch ops for the junk mask,
    CMC and AP (library kernels only; none of this repo's code).  Context for the headline, not
    a target and not the reference arm."""
    qp = torch.from_numpy(q_pid).to(qf.device).long()
    gp = torch.from_numpy(g_pid).to(qf.device).long()
    qc = torch.from_numpy(q_cam).to(qf.device).long()
    gc = torch.from_numpy(g_cam).to(qf.device).long()
    tf32 = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False

    def step():
        q = qf / torch.norm(qf, dim=1, keepdim=True)
        g = gf / torch.norm(gf, dim=1, keepdim=True)
        d = 1.0 - torch.mm(q, g.T)
        idx = torch.argsort(d, dim=1, stable=True)
        match = gp[idx] == qp[:, None]
        keep = ~(match & (gc[idx] == qc[:, None]))
        m = (match & keep).float()
        pos = torch.cumsum(keep.float(), dim=1)          # 1-based kept rank of every column
        hits = torch.cumsum(m, dim=1)
        nrel = m.sum(1)
        valid = nrel > 0
        ap = (hits / pos.clamp(min=1) * m).sum(1) / nrel.clamp(min=1)
        first = torch.where(m > 0, pos, torch.full_like(pos, float("inf"))).min(dim=1).values
        cmc1 = ((first <= 1) & valid).float().sum() / valid.float().sum()
        return float(ap[valid].mean().item()), float(cmc1.item())

    try:
        for _ in range(2):
            step()
        ms, (mAP, r1) = timed(step, 3)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = tf32
    Q, G = qf.shape[0], gf.shape[0]
    return {"what": "torch.mm (cuBLAS fp32) + torch.argsort + torch ops, device-resident features",
            "ms_per_step": ms / 3, "pairs_per_s": Q * G / (ms / 3 * 1e-3), "mAP": mAP, "rank1": r1}


def bind_to_gpu_numa_node(gpu_index):
    """Pin this process to the CPUs local to its GPU (NVML's ideal affinity) BEFORE any pinned host
    buffer is allocated, so that first-touch places the staging pages on the GPU's own NUMA node.
    With one rank per GPU on a two-socket host, ranks left unbound pull half of their H2D traffic
    across the socket interconnect.  Best effort: returns a short description or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            uuid = "GPU-" + str(torch.cuda.get_device_properties(gpu_index).uuid)
            h = pynvml.nvmlDeviceGetHandleByUUID(uuid)
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        ncpu = os.cpu_count() or 64
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        target = cpus & allowed
        if not target or target == allowed:
            return {"cpus_local_to_gpu": len(cpus), "bound": False}
        os.sched_setaffinity(0, target)
        return {"cpus_local_to_gpu": len(cpus), "bound": True, "cpus": len(target)}
    except Exception as e:  # pragma: no cover - best effort
        return {"bound": False, "error": str(e)[:80]}


def run_ours(args):
    import torch.distributed as dist
    from daliid_b200 import _lib, metrics, sharded

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: daliid_b200 has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 and not args.no_numa_bind else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    pk = peaks()
    wl = make_workload(WORKLOAD, rank, world, dev)
    cfg = wl["cfg"]
    Q, G, D = cfg["Q"], cfg["G"], cfg["D"]
    pairs_per_step = Q * G * world
    prec = args.precision
    ctx = _lib.get_ctx(local_rank)

    qf_d, gf_d = wl["qf"].to(dev), wl["gf"].to(dev)
    qf_h, gf_h = wl["qf"].pin_memory(), wl["gf"].pin_memory()
    g_pid = wl["g_pid_all"]
    g_cam = wl["g_cam_all"]

    def step(qf, gf):
        if world == 1:
            return metrics.evaluate_features(qf, gf, wl["q_pid"], g_pid, wl["q_cam"], g_cam,
                                             metric="cosine", precision=prec)
        return sharded.evaluate_features_sharded(qf, gf, wl["g0"], wl["q_pid"], g_pid, wl["q_cam"],
                                                 g_cam, metric="cosine", precision=prec)

    def step_host():
        # host (pinned) pointers: the library stages them -- chunked H2D overlapped with the
        # preparation + contraction of the previous chunk -- inside the call
        return step(qf_h, gf_h)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, out

    for _ in range(max(args.warmup, 3)):
        step(qf_d, gf_d)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # per-kernel CUDA events (they feed `roofline`) are recorded on every 4th step of the timed
    # region: a pair of events around each of the four launches costs ~0.03 ms per step (4 %), which
    # would otherwise sit in the headline; sampled steps are ordinary steps of the same loop
    sample_every = 4 if args.steps >= 8 else 1
    sampled = len(range(0, args.steps, sample_every))
    step_no = [0]

    def step_sampled():
        ctx.timing_enable(step_no[0] % sample_every == 0)
        step_no[0] += 1
        return step(qf_d, gf_d)

    ctx.timing_reset()
    n0 = ctx.launch_count()
    ms, (cmc, mAP) = timed(step_sampled, args.steps)
    launches = ctx.launch_count() - n0
    ctx.timing_enable(True)  # timing_read drains the pending events
    ktimes = ctx.timing_read()
    ctx.timing_enable(False)

    hits = ctx.plan_cache_hits()
    # the library times its host-input pipeline with one and two DMA streams during the first five
    # calls of a context and then keeps the faster setting: those calls are warm-up here
    for _ in range(max(args.warmup, 7)):
        step_host()
    ms_e2e, _ = timed(step_host, args.steps)
    hits = ctx.plan_cache_hits() - hits
    # the same end-to-end steps from PAGEABLE host memory: what the reference's producer hands over
    # (getFeatures.py:62-67 concatenates per-batch .cpu() tensors; nothing is pinned there)
    qf_p, gf_p = wl["qf"].clone(), wl["gf"].clone()
    for _ in range(3):
        step(qf_p, gf_p)
    ms_pageable, _ = timed(lambda: step(qf_p, gf_p), max(4, args.steps // 2))
    ms_pageable /= max(4, args.steps // 2)
    del qf_p, gf_p
    # host ceiling: the raw pinned copy of one step's input bytes, all ranks at once
    in_bytes = int((-(-Q // world) + G) * D * 4)
    h2d_ceiling = measure_h2d_ceiling(dev, in_bytes, world, dist)
    # H2D and peer-exchange shares of the end-to-end step (events on the copy / compute streams)
    ctx.timing_enable(True)
    ctx.timing_reset()
    for _ in range(4):
        step_host()
    kt_e2e = ctx.timing_read()
    ctx.timing_enable(False)
    # multi-GPU: the sharded result against the single-GPU evaluation of the WHOLE gallery
    sharded_parity = None
    if world > 1:
        slabs = [torch.empty_like(gf_d) for _ in range(world)]
        dist.all_gather(slabs, gf_d)
        g_all = torch.cat(slabs)
        del slabs
        s_cmc, s_map, s_det = sharded.evaluate_features_sharded(qf_d, gf_d, wl["g0"], wl["q_pid"], g_pid, wl["q_cam"],
                                                                g_cam, metric="cosine", precision=prec,
                                                                return_details=True)
        u_cmc, u_map, u_det = metrics.evaluate_features(qf_d, g_all, wl["q_pid"], g_pid, wl["q_cam"], g_cam,
                                                        metric="cosine", precision=prec, return_details=True)
        del g_all
        same = bool(np.array_equal(s_cmc, u_cmc) and s_map == u_map and
                    np.array_equal(s_det["first_rank"], u_det["first_rank"]) and
                    np.array_equal(s_det["ap"], u_det["ap"], equal_nan=True))
        flag = torch.tensor([1 if same else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        assert int(flag.item()) == 1, f"rank {rank}: sharded evaluation differs from the single-GPU one"
        sharded_parity = "bit-identical"
    # the same steps with the rank plan rebuilt from the labels on every step
    ctx.plan_cache_enable(False)
    step(qf_d, gf_d)
    ms_nocache, _ = timed(lambda: step(qf_d, gf_d), args.steps)
    ctx.plan_cache_enable(True)
    # clocks / throttle reasons sampled over all the timed loops above (the headline loop alone is
    # ~15 ms, shorter than one nvidia-smi sampling period)
    clocks = sampler.stop() if rank == 0 else None

    if args.breakdown:
        ops = sharded.CudaOps(local_rank)
        acc = {}

        def ph(name, fn):
            t0 = time.perf_counter()
            r = fn()
            t1 = time.perf_counter()
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            a = acc.setdefault(name, [0.0, 0.0])
            a[0] += t1 - t0
            a[1] += t2 - t0
            return r
        for it in range(10):
            d = ph("distmat", lambda: ops.distmat(qf_d, gf_d, "cosine", prec, True))
            qp_, gp_ = ph("labels", lambda: metrics.canonicalize_labels(wl["q_pid"], g_pid))
            qc_, gc_ = ph("labels", lambda: metrics.canonicalize_labels(wl["q_cam"], g_cam))
            plan = ph("plan", lambda: ops.plan(qp_, gp_, qc_, gc_))
            keys = ph("gather", lambda: ops.gather_keys(plan, d, wl["g0"]))
            ph("allreduce1", lambda: sharded._all_reduce_sum(keys, None))
            counts = ph("count", lambda: ops.count(plan, d, wl["g0"], keys))
            ph("allreduce2", lambda: sharded._all_reduce_sum(counts, None))
            ph("finalize", lambda: ops.finalize(plan, keys, counts, Q, G * world, 50, "cy_f32"))
            ops.plan_destroy(plan)
        # GPU-timeline version: no host syncs, CUDA events between the phases
        names = ["distmat", "plan", "gather", "allreduce1", "count", "allreduce2", "finalize"]
        tl = {n: 0.0 for n in names}
        host_total = 0.0
        for it in range(10):
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
            h0 = time.perf_counter()
            evs[0].record()
            d = ops.distmat(qf_d, gf_d, "cosine", prec, True); evs[1].record()
            qp_, gp_ = metrics.canonicalize_labels(wl["q_pid"], g_pid)
            qc_, gc_ = metrics.canonicalize_labels(wl["q_cam"], g_cam)
            plan = ops.plan(qp_, gp_, qc_, gc_); evs[2].record()
            keys = ops.gather_keys(plan, d, wl["g0"]); evs[3].record()
            sharded._all_reduce_sum(keys, None); evs[4].record()
            counts = ops.count(plan, d, wl["g0"], keys); evs[5].record()
            sharded._all_reduce_sum(counts, None); evs[6].record()
            h1 = time.perf_counter()
            ops.finalize(plan, keys, counts, Q, G * world, 50, "cy_f32"); evs[7].record()
            ops.plan_destroy(plan)
            torch.cuda.synchronize()
            host_total += h1 - h0
            for i, n in enumerate(names):
                tl[n] += evs[i].elapsed_time(evs[i + 1])
        if rank == 0:
            print("gpu timeline (ms/step): " + ", ".join(f"{k}={v / 10:.3f}" for k, v in tl.items()) +
                  f", total={sum(tl.values()) / 10:.3f}, host enqueue before finalize={host_total * 100:.3f}",
                  file=sys.stderr, flush=True)
            print("breakdown (ms: host call, call+sync): " +
                  ", ".join(f"{k}={v[0] * 100:.3f}/{v[1] * 100:.3f}" for k, v in acc.items()),
                  file=sys.stderr, flush=True)

    # the other arithmetic modes of the contraction on the same workload (BASELINE config 1:
    # "exact fp32 path vs TF32 tensor-core path"), a few steps each
    modes = {}
    if world == 1 and not args.no_modes:
        for m in ("fp32", "tf32", "f16", "tf32c"):
            def step_m(m=m):
                return metrics.evaluate_features(qf_d, gf_d, wl["q_pid"], g_pid, wl["q_cam"], g_cam,
                                                 metric="cosine", precision=m)
            for _ in range(2):
                step_m()
            ms_m, (_, map_m) = timed(step_m, 5)
            modes[m] = {"ms_per_step": ms_m / 5, "pairs_per_s": Q * G / (ms_m / 5 * 1e-3), "mAP": map_m}

    stock = None
    if world == 1 and rank == 0 and not args.no_modes:
        stock = stock_gpu_baseline(qf_d, gf_d, wl["q_pid"], g_pid, wl["q_cam"], g_cam, timed)

    c1 = c3 = c4 = c5 = None
    if not args.no_side:
        if world == 1:
            c1 = measure_c1(dev, timed, ctx)
            c4 = measure_c4(dev, timed, ctx)
        c3 = measure_c3(dev, rank, world, dist, timed, ctx)
    if not args.no_c5:
        tf32_peak = measure_tf32_peak(dev)
        c5 = measure_c5(dev, rank, world, dist, tf32_peak)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # roofline of the dominant kernel (the Q x G x D contraction), from live event timings
    n_dm, ms_dm = ktimes["distmat"]
    flops_per_launch = 2.0 * Q * G * D  # algorithmic: 2*D per pair, one slab per launch
    achieved = flops_per_launch / (ms_dm / max(n_dm, 1) * 1e-3) / 1e12 if n_dm else None
    n_rc, ms_rc = ktimes["rank_count"]
    rank_gbs = (4.0 * Q * G) / (ms_rc / max(n_rc, 1) * 1e-3) / 1e9 if n_rc else None
    roofline = {"bound": "tensor", "kernel": "distmat_umma2_kernel" if prec != "fp32" else "distmat_simt_kernel",
                "achieved": achieved, "peak": pk["bf16"], "unit": "TFLOP/s",
                "frac": (achieved / pk["bf16"]) if achieved else None,
                # dram__bytes_read.sum + dram__bytes_write.sum need ncu and are not measured in this run:
                # see profiles/ (r02_c2_full.md: read + write per launch of this kernel at this shape);
                # algorithmic minimum = 161 MB of fp16 operand planes + 214 MB of distance matrix
                "traffic": None,
                "traffic_source": "not measured live (ncu capture of the same command: profiles/r02b_c2_full.md)",
                "peak_source": pk["source"] + " bf16 burst; the fp32-class splits issue several tensor "
                               "passes per algorithmic FLOP: ceiling of frac = 1/3 for f16x3 (three 16-bit "
                               "passes), 1/2 for tf32, 1/4 for tf32c (1 TF32 + 2 bf16 passes), 1/6 for tf32x3",
                "avg_launch_ms": ms_dm / max(n_dm, 1) if n_dm else None}
    roofline_rank = {"bound": "hbm", "kernel": "rank_count_v3_kernel (one launch: thresholds, counting, CMC/AP epilogue)"
                     if world == 1 else "rank_count_v3_kernel", "achieved": rank_gbs,
                     "peak": pk["hbm"], "unit": "GB/s", "frac": (rank_gbs / pk["hbm"]) if rank_gbs else None,
                     "traffic": None,  # ncu only: profiles/r02b_c2_full.md
                     "avg_launch_ms": ms_rc / max(n_rc, 1) if n_rc else None}

    line = {
        "metric": METRIC, "value": pairs_per_step * args.steps / (ms * 1e-3), "unit": UNIT,
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32" if prec != "tf32" else "tf32",
        "data": "synthetic (seed 12, class centres + gaussian noise; no datasets offline)",
        "config": {"workload": f"{WORKLOAD}: Q={Q} x G={G}/GPU x D={D}, cosine, precision={prec}",
                   "global_gallery": G * world, "parallelism": f"gallery-sharded x{world}",
                   "weak_scaling": "one Market-sized slab per GPU; gallery identities = 751 x n_gpus "
                                   "(positives per query constant)",
                   "l2": "inputs larger than L2 (features 158 MB + operand planes 315 MB + distmat "
                         "214 MB per step vs 126 MB L2); no explicit flush"},
        "clocks": clocks,
        "e2e": {"value": pairs_per_step * args.steps / (ms_e2e * 1e-3), "unit": UNIT,
                "ms_per_step": ms_e2e / args.steps,
                # per rank: its share of the queries (the ranks exchange the shares over NVLink,
                # sharded.share_queries) + its gallery slab, fp32, + the int32 label arrays
                "h2d_bytes_per_step": int((-(-Q // world) + G) * D * 4 + (Q + G * world) * 8),
                "d2h_bytes_per_step": int(Q * 8 + 51 * 4), "h2d_dma_streams": ctx.h2d_streams()},
        "e2e_pageable": {"value": pairs_per_step / (ms_pageable * 1e-3), "unit": UNIT, "ms_per_step": ms_pageable,
                         "note": "same call, features in pageable host memory (what the reference's "
                                 "getFeatures.py hands over); the library stages them itself: worker threads "
                                 "copy 4 MB sub-chunks into pinned slots while the DMA engine drains the "
                                 "previous one (the CUDA driver's own pageable staging ran at 11 GB/s: 14.3 ms)"},
        "h2d_ceiling": h2d_ceiling,
        "e2e_kernel_ms_per_step": {k: v[1] / 4 for k, v in kt_e2e.items() if v[0]},
        "sharded_parity": sharded_parity,
        "gpu_launches": int(launches),
        "numa_binding": numa,
        "rank_plan": {"note": "the plan (gallery index by identity, label-only) of the previous step is "
                              "reused when the four label arrays compare equal byte for byte (checked "
                              "every step); ms_per_step_rebuilt = the same steps with the cache off",
                      "cache_hits_in_e2e_steps": int(hits), "ms_per_step_rebuilt": ms_nocache / args.steps},
        "roofline": roofline, "roofline_rank_stage": roofline_rank,
        "kernel_ms_per_step": {k: v[1] / sampled for k, v in ktimes.items() if v[0]},
        "kernel_events": "CUDA events around every launch of every %d%s step of the timed region" % (
            sample_every, {1: "st", 2: "nd", 3: "rd"}.get(sample_every, "th")),
        "mAP": mAP, "rank1": float(cmc[0]),
    }
    if modes:
        line["other_precisions"] = modes
    if stock:
        line["stock_gpu_baseline"] = stock
    if c1:
        line["c1_market_vit"] = c1
    if c3:
        line["c3_deepchange"] = c3
    if c4:
        line["c4_fusion3"] = c4
    if c5:
        line["c5_faceid_1toN"] = c5
    if world == 1 and not args.no_cpu_baseline:
        from oracle import c_oracle
        c_oracle.build()
        a = (wl["qf"], wl["gf"], wl["q_pid"], g_pid[:G], wl["q_cam"], g_cam[:G])
        t0 = time.perf_counter()
        c_cmc, c_map = cpu_reference_step(*a)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": Q * G / dt, "unit": UNIT, "cores": torch.get_num_threads(),
                                "kind": "port", "seconds": dt,
                                "sample": "the full workload once (CPU torch.mm on all threads + np.argsort "
                                          "+ compiled evaluator loop, single-threaded as upstream)",
                                "mAP": c_map}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-numa-bind", action="store_true",
                    help="multi-GPU runs: leave the ranks' CPU affinity alone (default: bind each rank to its GPU's NUMA node)")
    ap.add_argument("--precision", default="f16x3", choices=["fp32", "tf32x3", "tf32c", "tf32", "f16x3", "f16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-modes", action="store_true", help="skip the per-precision side measurements")
    ap.add_argument("--no-c5", action="store_true", help="skip the config-5 (1:N top-k) side measurement")
    ap.add_argument("--no-side", action="store_true", help="skip the C1 / C3 / C4 side measurements")
    ap.add_argument("--breakdown", action="store_true", help="diagnostic per-phase host timing (stderr)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

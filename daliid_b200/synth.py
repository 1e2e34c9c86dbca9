"""Synthetic inputs of the shapes BASELINE.json names.

There are no datasets or trained weights in this environment, so features are drawn as
``centers[pid] + sigma * N(0,1)`` with the reference's seed 12 (``evaluate.py:49-50``),
which gives non-degenerate retrieval quality (mAP well inside (0,1))."""
from __future__ import annotations

import numpy as np
import torch

CONFIGS = {
    # name: Q, G, D, n_ids, n_cams, sigma
    "market_vit": dict(Q=3368, G=15913, D=768, n_ids=751, n_cams=6, sigma=2.5),
    "market_resnet50": dict(Q=3368, G=15913, D=2048, n_ids=751, n_cams=6, sigma=4.0),
    "deepchange": dict(Q=17527, G=62956, D=768, n_ids=521, n_cams=3400, sigma=3.0),
    "tiny": dict(Q=97, G=403, D=72, n_ids=23, n_cams=4, sigma=1.5),
    "small": dict(Q=300, G=2100, D=200, n_ids=61, n_cams=5, sigma=2.0),
}


def make_labels(Q, G, n_ids, n_cams, seed=12):
    g = torch.Generator().manual_seed(seed)
    g_pid = torch.randint(0, n_ids, (G,), generator=g)
    g_cam = torch.randint(0, n_cams, (G,), generator=g)
    q_pid = torch.randint(0, n_ids, (Q,), generator=g)
    q_cam = torch.randint(0, n_cams, (Q,), generator=g)
    return (q_pid.numpy().astype(np.int32), g_pid.numpy().astype(np.int32),
            q_cam.numpy().astype(np.int32), g_cam.numpy().astype(np.int32))


def make_features(Q, G, D, n_ids, n_cams, sigma, seed=12, device="cpu"):
    """Returns (qf [Q,D], gf [G,D], q_pid, g_pid, q_cam, g_cam); features fp32 on ``device``."""
    q_pid, g_pid, q_cam, g_cam = make_labels(Q, G, n_ids, n_cams, seed)
    dev = torch.device(device)
    g = torch.Generator(device=dev).manual_seed(seed)
    centers = torch.randn(n_ids, D, generator=g, device=dev)
    qf = centers[torch.as_tensor(q_pid, device=dev).long()] + sigma * torch.randn(Q, D, generator=g, device=dev)
    gf = centers[torch.as_tensor(g_pid, device=dev).long()] + sigma * torch.randn(G, D, generator=g, device=dev)
    return qf.contiguous(), gf.contiguous(), q_pid, g_pid, q_cam, g_cam


def make_config(name, seed=12, device="cpu", **override):
    cfg = dict(CONFIGS[name])
    cfg.update(override)
    return make_features(seed=seed, device=device, **cfg)


def as_reference_rows(pid, cam, kind="person"):
    """Label arrays -> the reference's ``ndarray[str] [N,4] = [path, pid, camid, kind]`` rows
    (``datasetUtils.py:15-17``), which is what the drop-in entry points receive."""
    n = len(pid)
    return np.array([["img_%06d.jpg" % i, str(int(pid[i])), str(int(cam[i])), kind] for i in range(n)])

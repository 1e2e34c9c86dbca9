"""Generates the golden fixtures under tests/golden/ (run in the BUILD container, where
/root/reference exists; the fixtures then travel to the GPU box, the reference does not).

Two kinds of expectation are recorded:

* outputs of the REFERENCE ITSELF run here: the distance / fusion / top-k expressions are
  plain torch/numpy and are executed verbatim (via oracle.distmat_oracle, which restates
  them line by line), and the reference's own call-site functions
  (``validateModels.calculateMetrics``, ``validateBRIAR.calculateMetrics``,
  ``evaluate.calculate_metrics``, ``evaluateCleanATModels.calculateMetrics``) are lifted
  from /root/reference/Person-ReID/*.py by AST and executed unmodified, with the one
  missing third-party symbol (``torchreid.metrics.evaluate_rank``) bound to the oracle;
* outputs of the restated oracle for the third-party evaluator (PARITY UNPINNED: the
  reference holds no fixture for it).

    python tests/golden/make_golden.py
"""
from __future__ import annotations

import ast
import contextlib
import io
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import distmat_oracle as do  # noqa: E402
from oracle import rank_oracle as ro  # noqa: E402

REF = "/root/reference/Person-ReID"


def lift(path, class_name, func_name):
    """Return the source of one function/method of the reference, unmodified."""
    src = open(path).read()
    tree = ast.parse(src)
    nodes = tree.body
    if class_name:
        cls = [n for n in nodes if isinstance(n, ast.ClassDef) and n.name == class_name][0]
        nodes = cls.body
    fn = [n for n in nodes if isinstance(n, ast.FunctionDef) and n.name == func_name][0]
    lines = src.splitlines()[fn.lineno - 1: fn.end_lineno]
    import textwrap
    return textwrap.dedent("\n".join(lines).expandtabs(4)), fn.lineno, fn.end_lineno


def run_reference_callsite(path, class_name, func_name, args, accum):
    """Execute a lifted reference function with torchreid.metrics.evaluate_rank -> oracle."""
    code, lo, hi = lift(path, class_name, func_name)
    torchreid = types.SimpleNamespace(metrics=types.SimpleNamespace(
        evaluate_rank=lambda *a, **k: ro.evaluate_rank(*a, accum=accum, **k)))
    ns = {"torch": torch, "np": np, "torchreid": torchreid}
    exec(compile(code, f"{path}:{lo}-{hi}", "exec"), ns)
    fn = ns[func_name]
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        ret = fn(*args) if not class_name else fn(types.SimpleNamespace(), *args)
    return ret, buf.getvalue(), f"{os.path.basename(path)}:{lo}-{hi}"


def rows(pid, cam):
    return np.array([["img_%05d.jpg" % i, str(int(pid[i])), str(int(cam[i])), "person"]
                     for i in range(len(pid))])


def rank_expect(dist, qp, gp, qc, gc, max_rank=50):
    out = {}
    for accum in ("cy_f32", "py_f64"):
        fn = {"cy_f32": ro.eval_market1501_cy_f32, "py_f64": ro.eval_market1501_py_f64}[accum]
        cmc, mAP, ap, first = fn(dist, qp, gp, qc, gc, max_rank, return_details=True)
        out[f"cmc_{accum}"] = cmc
        out[f"mAP_{accum}"] = np.float64(mAP)
        out[f"ap_{accum}"] = ap
        out["first_rank"] = first.astype(np.int32)
    return out


def main():
    rng = np.random.default_rng(12)
    torch.manual_seed(12)
    meta = {}

    # ---- case "tiny": features -> reference distance expressions -> oracle metrics ------
    Q, G, D, n_ids, n_cams = 97, 403, 72, 23, 4
    g_pid = rng.integers(0, n_ids, G); g_cam = rng.integers(0, n_cams, G)
    q_pid = rng.integers(0, n_ids + 2, Q); q_cam = rng.integers(0, n_cams, Q)  # ids 23,24 absent
    centers = rng.standard_normal((n_ids + 2, D)).astype(np.float32)
    qf = (centers[q_pid] + 1.5 * rng.standard_normal((Q, D))).astype(np.float32)
    gf = (centers[g_pid] + 1.5 * rng.standard_normal((G, D))).astype(np.float32)
    tq, tg = torch.from_numpy(qf), torch.from_numpy(gf)
    cos = do.cosine_distmat(tq, tg).numpy()
    tiny = dict(qf=qf, gf=gf, q_pid=q_pid, g_pid=g_pid, q_cam=q_cam, g_cam=g_cam,
                qn=do.l2_normalize(tq).numpy(), q_norm=torch.norm(tq, dim=1).numpy(),
                cosine=cos, sqeuclidean=do.sqeuclidean_distmat(tq, tg).numpy(),
                euclidean=do.euclidean_distmat(tq, tg).numpy(), dot=torch.mm(tq, tg.T).numpy())
    tiny.update(rank_expect(cos, q_pid, g_pid, q_cam, g_cam))
    np.savez_compressed(os.path.join(HERE, "tiny.npz"), **tiny)

    # reference call sites on the tiny case (string label rows, as the loaders build them)
    queries, gallery = rows(q_pid, q_cam), rows(g_pid, g_cam)
    calls = {}
    for accum in ("cy_f32", "py_f64"):
        (cmc, mAP), out, where = run_reference_callsite(
            f"{REF}/validateModels.py", "validateModels", "calculateMetrics",
            (torch.from_numpy(cos), queries, gallery), accum)
        calls[f"validateModels.calculateMetrics[{accum}]"] = dict(
            where=where, stdout=out, mAP=float(mAP), cmc=[float(x) for x in cmc])
        _, out, where = run_reference_callsite(
            f"{REF}/evaluate.py", None, "calculate_metrics", (cos, queries, gallery), accum)
        calls[f"evaluate.calculate_metrics[{accum}]"] = dict(where=where, stdout=out)
        _, out, where = run_reference_callsite(
            f"{REF}/evaluateCleanATModels.py", None, "calculateMetrics", (queries, gallery, cos), accum)
        calls[f"evaluateCleanATModels.calculateMetrics[{accum}]"] = dict(where=where, stdout=out)
    # validateBRIAR: needs tie-free top-20 (torch.argsort is not stable by default)
    st = torch.argsort(torch.from_numpy(cos), dim=1, stable=True)[:, :20]
    assert torch.equal(st, torch.argsort(torch.from_numpy(cos), dim=1)[:, :20])
    (cmc_b, zero), out, where = run_reference_callsite(
        f"{REF}/validateModels.py", "validateBRIAR", "calculateMetrics",
        (torch.from_numpy(cos), queries, gallery), "cy_f32")
    calls["validateBRIAR.calculateMetrics"] = dict(where=where, stdout=out,
                                                   cmc=[float(x) for x in cmc_b], second=int(zero))
    meta["callsites"] = calls
    np.save(os.path.join(HERE, "tiny_top20.npy"), st.numpy().astype(np.int32))

    # ---- case "ties": adversarial distance matrix for the rank stage ---------------------
    Q, G = 64, 500
    d = (np.round(rng.random((Q, G)) * 64) / 64).astype(np.float32)   # heavy exact ties
    d[:, 100:140] = d[:, 60:100]                                       # duplicated columns
    d[3, :] = np.nan                                                   # zero-norm query row
    d[4, ::7] = np.nan
    d[5, ::5] = np.inf
    d[6, ::3] = -0.0; d[6, 1::3] = 0.0
    d[7, :] = 0.25                                                     # a fully tied row
    d[8, ::2] = -np.inf
    g_pid = rng.integers(0, 12, G); g_cam = rng.integers(0, 3, G)
    q_pid = rng.integers(0, 14, Q); q_cam = rng.integers(0, 3, Q)     # ids 12,13 absent
    # a query whose only matches are junk (same pid, same cam) -> invalid
    g_pid[g_pid == 11] = 10
    g_pid[0] = 11; g_cam[0] = 2; q_pid[9] = 11; q_cam[9] = 2
    ties = dict(dist=d, q_pid=q_pid, g_pid=g_pid, q_cam=q_cam, g_cam=g_cam)
    ties.update(rank_expect(d, q_pid, g_pid, q_cam, g_cam))
    ties["top20"] = ro.stable_argsort(d)[:, :20].astype(np.int32)
    np.savez_compressed(os.path.join(HERE, "ties.npz"), **ties)

    # ---- case "small_gallery": G < max_rank clamps max_rank -------------------------------
    Q, G = 12, 17
    d = rng.random((Q, G)).astype(np.float32)
    g_pid = rng.integers(0, 4, G); g_cam = rng.integers(0, 2, G)
    q_pid = rng.integers(0, 4, Q); q_cam = rng.integers(0, 2, Q)
    sg = dict(dist=d, q_pid=q_pid, g_pid=g_pid, q_cam=q_cam, g_cam=g_cam)
    sg.update(rank_expect(d, q_pid, g_pid, q_cam, g_cam))
    np.savez_compressed(os.path.join(HERE, "small_gallery.npz"), **sg)

    # ---- case "fusion": reference fusion expressions -----------------------------------
    Q, G = 53, 211
    ds = [rng.random((Q, G)).astype(np.float32) * 2 for _ in range(3)]
    qm = [torch.from_numpy((rng.random((Q, 1)) * 30 + 1).astype(np.float32)) for _ in range(2)]
    gm = [torch.from_numpy((rng.random((G, 1)) * 30 + 1).astype(np.float32)) for _ in range(2)]
    w = [do.magnitude_weights(qm[i], gm[i]) for i in range(2)]
    fusion = dict(d0=ds[0], d1=ds[1], d2=ds[2],
                  mean2=do.fuse_mean(ds[:2]), mean3=do.fuse_mean(ds),
                  qm0=qm[0].numpy(), qm1=qm[1].numpy(), gm0=gm[0].numpy(), gm1=gm[1].numpy(),
                  weighted=do.fuse_weighted(w, ds[:2]).numpy())
    assert fusion["mean2"].dtype == np.float32 and fusion["weighted"].dtype == np.float32
    np.savez_compressed(os.path.join(HERE, "fusion.npz"), **fusion)

    # ---- hand-derived known-answer cases --------------------------------------------------
    meta["kat"] = [
        dict(name="two_junk_two_positives",
             dist=[[0.1, 0.2, 0.3, 0.4, 0.5, 0.6]], q_pid=[7], q_cam=[0],
             g_pid=[7, 1, 7, 7, 2, 7], g_cam=[0, 1, 1, 0, 1, 2],
             mAP=0.5, cmc=[0, 1, 1, 1, 1, 1], first_rank=[2]),
        dict(name="first_hit_rank1_and_invalid_query",
             dist=[[0.3, 0.1, 0.2], [0.1, 0.2, 0.3]], q_pid=[5, 9], q_cam=[0, 0],
             g_pid=[5, 5, 6], g_cam=[1, 1, 1],
             # order 1(+),2(-),0(+): AP=(1/1+2/3)/2; second query has no match -> skipped
             mAP=5.0 / 6.0, cmc=[1, 1, 1], first_rank=[1, -1]),
        dict(name="tie_broken_by_index",
             dist=[[0.5, 0.5, 0.5, 0.5]], q_pid=[1], q_cam=[0],
             g_pid=[2, 1, 2, 1], g_cam=[1, 1, 1, 1],
             # stable order 0,1,2,3 -> positives at ranks 2 and 4 -> AP=(1/2+2/4)/2
             mAP=0.5, cmc=[0, 1, 1, 1], first_rank=[2]),
    ]
    with open(os.path.join(HERE, "meta.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    for fn in sorted(os.listdir(HERE)):
        print(fn, os.path.getsize(os.path.join(HERE, fn)))


if __name__ == "__main__":
    main()

"""ctypes binding of ``rank_oracle.c`` (CPU ORACLE -- test infrastructure only).

``evaluate_rank_c`` mirrors what an installed torchreid does when its Cython
extension is present (the reference probes for that at ``validateModels.py:16-24``):
numpy does the argsort, compiled code does the per-query loop.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

from .rank_oracle import NO_VALID_MSG, canonicalize_labels

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "librank_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "rank_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
        i64p = ctypes.POINTER(ctypes.c_int64)
        f32p = ctypes.POINTER(ctypes.c_float)
        _lib.oracle_eval_market1501_cy.restype = ctypes.c_int
        _lib.oracle_eval_market1501_cy.argtypes = [
            i64p, i64p, i64p, i64p, i64p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
            f32p, f32p, f32p, i64p, i64p]
        _lib.oracle_stable_argsort_rows.restype = None
        _lib.oracle_stable_argsort_rows.argtypes = [
            f32p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, i64p]
    return _lib


def _p(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


def evaluate_rank_c(distmat, q_pids, g_pids, q_camids, g_camids, max_rank=50,
                    tie="stable", return_details=False):
    distmat = np.ascontiguousarray(np.asarray(distmat), dtype=np.float32)
    num_q, num_g = distmat.shape
    qp, gp = canonicalize_labels(q_pids, g_pids)
    qc, gc = canonicalize_labels(q_camids, g_camids)
    if num_g < max_rank:
        max_rank = num_g
    if tie == "stable":
        indices = np.argsort(distmat, axis=1, kind="stable")
    elif tie == "numpy_default":
        indices = np.argsort(distmat, axis=1)  # upstream's literal call
    elif tie == "c_stable":
        indices = np.empty((num_q, num_g), dtype=np.int64)
        lib().oracle_stable_argsort_rows(_p(distmat, ctypes.c_float), num_q, num_g, num_g,
                                         _p(indices, ctypes.c_int64))
    else:
        raise ValueError(tie)
    indices = np.ascontiguousarray(indices, dtype=np.int64)
    cmc = np.zeros(max_rank, dtype=np.float32)
    mAP = ctypes.c_float(0.0)
    ap = np.zeros(num_q, dtype=np.float32)
    fr = np.zeros(num_q, dtype=np.int64)
    nv = ctypes.c_int64(0)
    qp, gp, qc, gc = (np.ascontiguousarray(x, dtype=np.int64) for x in (qp, gp, qc, gc))
    rc = lib().oracle_eval_market1501_cy(
        _p(indices, ctypes.c_int64), _p(qp, ctypes.c_int64), _p(gp, ctypes.c_int64),
        _p(qc, ctypes.c_int64), _p(gc, ctypes.c_int64), num_q, num_g, max_rank,
        _p(cmc, ctypes.c_float), ctypes.byref(mAP), _p(ap, ctypes.c_float),
        _p(fr, ctypes.c_int64), ctypes.byref(nv))
    assert rc == 0, NO_VALID_MSG
    if return_details:
        apd = ap.astype(np.float64)
        apd[fr < 0] = np.nan
        return cmc, float(mAP.value), apd, fr
    return cmc, float(mAP.value)

"""Operator-level boundary of the retrieval-evaluation hot path.

Every function here is a thin host-side mirror of a reference call, backed by the C-ABI
library (``include/daliid_b200.h``).  No arithmetic of the hot path runs in Python.

==============================  ====================================================
this module                     reference call it replaces
==============================  ====================================================
``evaluate_rank``               ``torchreid.metrics.evaluate_rank`` called at
                                validateModels.py:68-69, evaluate.py:312-313,
                                evaluate_ensembled_models.py:324-325,
                                evaluateCleanATModels.py:266-267
``compute_distance_matrix``     ``1.0 - torch.mm(q, g.T)`` validateModels.py:47,
                                evaluate.py:291 (+ the commented euclidean calls
                                validateModels.py:44-45)
``normalize``                   ``x/torch.norm(x, dim=1, keepdim=True)``
                                validateModels.py:41-42
``fuse_distmats``               evaluate.py:278, evaluate_ensembled_models.py:313,
                                evaluateCleanATModels.py:127,154-157
``topk_identify``               ``torch.argsort(distmat, dim=1)[:, :20]``
                                validateModels.py:93; ``torch.topk`` :180
``evaluate_features``           the whole of validateModels.validate's arithmetic
                                (validateModels.py:41-47,61-69), fused
==============================  ====================================================
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from ._lib import ACCUMS, METRICS, PRECISIONS, as_i32, as_matrix, c_vp, get_ctx, np_ptr, p_i32

__all__ = [
    "re_ranking", "mrfuse", "argsort_rows",
    "canonicalize_labels",
    "evaluate_rank",
    "evaluate_rank_detailed",
    "compute_distance_matrix",
    "normalize",
    "fuse_distmats",
    "topk_identify",
    "topk_features",
    "evaluate_features",
]

DEFAULT_PRECISION = "auto"  # fp32-class result on the tensor pipe, see _precision()


F16X3_MAX_D = 4096  # above it the accumulator-truncation bound of f16x3 exceeds 1e-5


def _precision(precision, normalize, D=None):
    """"auto" = the fastest arithmetic that keeps distances within 1e-5 of the reference's fp32
    expression: the fp16 hi/lo split (f16x3) for unit rows of up to 4096 elements -- the reference's
    live cosine path at every feature size it has (512 .. 3840) -- the exact FP32 pipe for longer
    unit rows, and TF32 + bf16 corrections (tf32c) for un-normalised features."""
    if precision == "auto":
        if normalize:
            precision = "f16x3" if (D is None or D <= F16X3_MAX_D) else "fp32"
        else:
            precision = "tf32c"
    return _enum(PRECISIONS, precision, "precision")


def canonicalize_labels(q, g):
    """str / int label columns -> dense int32 ids (only equality matters).

    The reference's label columns are numpy *strings* (``datasetUtils.py:15-17``); they are
    passed as ``queries[:,1]``, ``gallery[:,1]`` (pid) and ``[:,2]`` (camid)."""
    q = np.asarray(q).reshape(-1)
    g = np.asarray(g).reshape(-1)
    if q.dtype == np.int32 and g.dtype == np.int32:
        return np.ascontiguousarray(q), np.ascontiguousarray(g)
    if q.dtype.kind in "iu" and g.dtype.kind in "iu" and q.size + g.size > 0:
        lo = min(q.min(initial=0), g.min(initial=0))
        hi = max(q.max(initial=0), g.max(initial=0))
        if lo >= np.iinfo(np.int32).min and hi <= np.iinfo(np.int32).max:
            return q.astype(np.int32), g.astype(np.int32)
    both = np.concatenate([q, g])
    _, inv = np.unique(both, return_inverse=True)
    inv = inv.astype(np.int32)
    return np.ascontiguousarray(inv[: q.shape[0]]), np.ascontiguousarray(inv[q.shape[0]:])


def _device_of(*bufs):
    for b in bufs:
        if b is not None and b.device is not None:
            return b.device
    return None


def _ctx_for(*bufs):
    ctx = get_ctx(_device_of(*bufs))
    ctx.attach_torch_stream()
    return ctx


def _enum(table, key, what):
    if isinstance(key, int):
        return key
    try:
        return table[key]
    except KeyError:
        raise ValueError(f"unknown {what} {key!r}; expected one of {sorted(table)}") from None


def evaluate_rank_detailed(distmat, q_pids, g_pids, q_camids, g_camids, max_rank=50,
                           accum="cy_f32"):
    """``evaluate_rank`` plus per-query AP, first-match rank and the number of valid queries."""
    d = as_matrix(distmat, np.float32, "distmat")
    num_q, num_g = d.shape
    qp, gp = canonicalize_labels(q_pids, g_pids)
    qc, gc = canonicalize_labels(q_camids, g_camids)
    if qp.shape[0] != num_q or gp.shape[0] != num_g or qc.shape[0] != num_q or gc.shape[0] != num_g:
        raise ValueError("label arrays do not match the distance matrix shape")
    if num_g < max_rank:
        max_rank = num_g
        print("Note: number of gallery samples is quite small, got {}".format(num_g))
    if num_q == 0 or num_g == 0:
        raise AssertionError(_lib.NO_VALID_MSG)
    ctx = _ctx_for(d)
    cmc = np.zeros(max_rank, dtype=np.float32)
    mAP = ctypes.c_double(0.0)
    ap = np.zeros(num_q, dtype=np.float64)
    first = np.zeros(num_q, dtype=np.int32)
    nvalid = ctypes.c_int64(0)
    rc = ctx.lib.dali_eval_rank_f32(
        ctx.h, c_vp(d.ptr), num_q, num_g, d.ld, p_i32(qp), p_i32(gp), p_i32(qc), p_i32(gc),
        int(max_rank), _enum(ACCUMS, accum, "accumulation mode"),
        np_ptr(cmc), ctypes.byref(mAP), np_ptr(ap), np_ptr(first), ctypes.byref(nvalid))
    ctx.check(rc)
    return cmc, float(mAP.value), ap, first, int(nvalid.value)


def evaluate_rank(distmat, q_pids, g_pids, q_camids, g_camids, max_rank=50,
                  use_metric_cuhk03=False, use_cython=True):
    """Drop-in for ``torchreid.metrics.evaluate_rank`` (market1501 protocol).

    Returns ``(cmc float32[max_rank], mAP float)``.  ``use_cython`` selects which of the two
    upstream accumulation semantics is reproduced bit for bit (True: compiled float32 path,
    False: pure-Python float64 path).  Ties are broken like a stable argsort.
    Raises ``AssertionError`` when no query identity appears in the gallery."""
    if use_metric_cuhk03:
        raise NotImplementedError("the cuhk03 protocol is never used by the reference")
    cmc, mAP, _, _, _ = evaluate_rank_detailed(
        distmat, q_pids, g_pids, q_camids, g_camids, max_rank,
        accum="cy_f32" if use_cython else "py_f64")
    return cmc, mAP


def _alloc_out(shape, like_device, dtype="float32"):
    """Output on the same side as the inputs: torch CUDA tensor or numpy array."""
    if like_device is not None:
        import torch
        t = torch.empty(shape, dtype=getattr(torch, dtype), device=f"cuda:{like_device}")
        return t, t.data_ptr()
    a = np.empty(shape, dtype=dtype)
    return a, a.ctypes.data


def normalize(x, return_norms=False):
    """Row L2 normalisation, no eps (validateModels.py:41-42).  Output lives where ``x`` does."""
    xb = as_matrix(x, np.float32, "x")
    n, d = xb.shape
    ctx = _ctx_for(xb)
    out, optr = _alloc_out((n, d), xb.device)
    norms, nptr = (_alloc_out((n,), xb.device) if return_norms else (None, None))
    if n:
        ctx.check(ctx.lib.dali_normalize_f32(ctx.h, c_vp(xb.ptr), n, d, xb.ld, c_vp(optr), d,
                                             c_vp(nptr)))
    return (out, norms) if return_norms else out


def compute_distance_matrix(input1, input2, metric="cosine", precision=DEFAULT_PRECISION,
                            normalize=None, out=None, device=None, padded=None):
    """``[Q,G]`` fp32 distance matrix between feature rows.

    ``metric``: ``cosine`` (``1 - q.g`` on L2-normalised rows), ``sqeuclidean`` (what
    torchreid's ``compute_distance_matrix(.., "euclidean")`` returns), ``euclidean``
    (``torch.cdist``) or ``dot``.  ``normalize`` defaults to True for cosine only.
    The result lives where the inputs live (CUDA tensor in, CUDA tensor out) unless ``out``
    (a contiguous float32 CUDA tensor or numpy array) or ``device`` (a CUDA ordinal: host
    features are streamed to the GPU in chunks overlapped with compute, the matrix stays
    there) says otherwise.  CUDA results are by default (``padded=None``) a ``[Q,G]`` view of a
    matrix whose rows are padded to a multiple of 4 floats: 16-byte aligned rows let the
    contraction store its tiles through TMA (a contiguous layout with an odd G, e.g. 15913,
    cannot).  ``padded=False`` returns a contiguous tensor; numpy results always are."""
    a = as_matrix(input1, np.float32, "input1")
    b = as_matrix(input2, np.float32, "input2")
    if a.shape[1] != b.shape[1]:
        raise ValueError("feature dimensions differ")
    if a.ld != a.shape[1] or b.ld != b.shape[1]:
        raise ValueError("feature matrices must be contiguous")
    Q, D = a.shape
    G = b.shape[0]
    m = _enum(METRICS, metric, "metric")
    if normalize is None:
        normalize = m == METRICS["cosine"]
    dev = _device_of(a, b)
    if out is not None:
        ob = as_matrix(out, np.float32, "out")
        own = out.data_ptr() if hasattr(out, "data_ptr") else np.asarray(out).ctypes.data
        if ob.ptr != own or ob.shape != (Q, G):
            raise ValueError("out must be a [Q,G] float32 array with unit column stride")
        optr, ld = ob.ptr, ob.ld
        if dev is None:
            dev = ob.device
    else:
        if dev is None and device is not None:
            dev = int(device)
        if (padded or padded is None) and dev is not None and G % 4 and Q:
            ld = (G + 3) // 4 * 4
            buf, optr = _alloc_out((Q, ld), dev)
            out = buf[:, :G]
        else:
            out, optr = _alloc_out((Q, G), dev)
            ld = G
    ctx = get_ctx(dev)
    ctx.attach_torch_stream()
    if Q and G:
        ctx.check(ctx.lib.dali_distmat_f32(ctx.h, c_vp(a.ptr), Q, c_vp(b.ptr), G, D, m,
                                           _precision(precision, normalize, D),
                                           1 if normalize else 0, c_vp(optr), max(ld, 1)))
    return out


def ensemble_distance_matrices(q_feats, g_feats, metric="cosine", precision=DEFAULT_PRECISION,
                               normalize=None, individual=True, device=None):
    """The distance matrices of an N-model ensemble and their mean ``((d0+d1)+..)/N`` in one sweep:
    every model's contraction adds its tile to the running sum in its epilogue, so the separate
    fusion pass over N+1 matrices (evaluate.py:278, evaluate_ensembled_models.py:313) disappears.

    Returns ``(distmats, mean)``; ``distmats`` is ``None`` with ``individual=False`` (then only the
    mean is written).  The mean is bit-identical to ``fuse_distmats`` of the separate matrices.
    Shapes the fused epilogue does not take (FP32-pipe precision, host-resident results) go through
    ``compute_distance_matrix`` + ``fuse_distmats`` -- same values."""
    n = len(q_feats)
    if n < 1 or n > 8 or len(g_feats) != n:
        raise ValueError("between 1 and 8 (query, gallery) feature pairs")
    a = [as_matrix(x, np.float32, "q_feats") for x in q_feats]
    b = [as_matrix(x, np.float32, "g_feats") for x in g_feats]
    Q, G = a[0].shape[0], b[0].shape[0]
    for x, y in zip(a, b):
        if x.shape[0] != Q or y.shape[0] != G or x.shape[1] != y.shape[1]:
            raise ValueError("all models must share the query and gallery sets")
        if x.ld != x.shape[1] or y.ld != y.shape[1]:
            raise ValueError("feature matrices must be contiguous")
    m = _enum(METRICS, metric, "metric")
    if normalize is None:
        normalize = m == METRICS["cosine"]
    dev = _device_of(*a, *b)
    if dev is None and device is not None:
        dev = int(device)

    def unfused():
        ds = [compute_distance_matrix(x, y, metric, precision, normalize, device=device)
              for x, y in zip(q_feats, g_feats)]
        return (ds if individual else None), fuse_distmats(ds)
    if dev is None or not Q or not G:
        return unfused()
    ctx = get_ctx(dev)
    ctx.attach_torch_stream()
    ld = (G + 3) // 4 * 4
    lda = (G + 7) // 8 * 8  # the running sum is re-read in 32-byte pieces
    acc_buf, acc_ptr = _alloc_out((Q, lda), dev)
    outs = []
    for i, (x, y) in enumerate(zip(a, b)):
        if individual:
            o_buf, o_ptr = _alloc_out((Q, ld), dev)
            outs.append(o_buf[:, :G])
        else:
            o_ptr = None
        rc = ctx.lib.dali_distmat_fuse_mean_f32(ctx.h, c_vp(x.ptr), Q, c_vp(y.ptr), G, x.shape[1], m,
                                                _precision(precision, normalize, x.shape[1]),
                                                1 if normalize else 0, c_vp(o_ptr), ld, c_vp(acc_ptr), lda,
                                                i, n)
        if rc == _lib.ERR_UNSUPPORTED:
            return unfused()
        ctx.check(rc)
    return (outs if individual else None), acc_buf[:, :G]


def fuse_distmats(distmats, q_weights=None, g_weights=None):
    """Fuse N distance matrices.

    Without weights: ``((d0+d1)+..)/N`` (evaluate.py:278, evaluate_ensembled_models.py:313).
    With ``q_weights[m]`` ([Q]) and ``g_weights[m]`` ([G]) -- the feature magnitudes of
    evaluateCleanATModels.py:249-256 -- ``w_m = max(qw_m[i], gw_m[j])`` and the result is
    ``sum(w_m d_m)/sum(w_m)`` (evaluateCleanATModels.py:154-157).  Bit-identical to the
    reference's numpy / torch expression."""
    bufs = [as_matrix(d, np.float32, "distmat") for d in distmats]
    if len({b.ld for b in bufs}) > 1 and all(b.device is not None for b in bufs):
        # mixed row pitches (a padded view next to its .clone()): bring all to the contiguous layout
        bufs = [as_matrix(b.keep.contiguous(), np.float32, "distmat") for b in bufs]
    n = len(bufs)
    if n < 1 or n > 8:
        raise ValueError("between 1 and 8 matrices can be fused")
    shape = bufs[0].shape
    dev = bufs[0].device
    for b in bufs:
        if b.shape != shape or b.device != dev or b.ld != bufs[0].ld:
            raise ValueError("all matrices must share shape, layout and device")
    Q, G = shape
    ctx = _ctx_for(*bufs)
    ld = bufs[0].ld if Q > 1 else G
    if ld != G:
        # row-padded device matrices (what compute_distance_matrix returns on CUDA): same layout out
        if dev is None:
            raise ValueError("host distance matrices must be contiguous")
        buf, optr = _alloc_out((Q, ld), dev)
        out = buf[:, :G]
    else:
        out, optr = _alloc_out((Q, G), dev)
    dptr = (c_vp * n)(*[b.ptr for b in bufs])
    wq = wg = None
    keep = []
    if (q_weights is None) != (g_weights is None):
        raise ValueError("q_weights and g_weights go together")
    if q_weights is not None:
        if len(q_weights) != n or len(g_weights) != n:
            raise ValueError("one weight vector pair per matrix")

        def vec(v, length):
            vb = as_matrix(v.reshape(1, -1) if hasattr(v, "reshape") else np.asarray(v).reshape(1, -1),
                           np.float32, "weights")
            if vb.shape[1] != length:
                raise ValueError("weight vector length mismatch")
            keep.append(vb)
            return vb.ptr

        wq = (c_vp * n)(*[vec(v, Q) for v in q_weights])
        wg = (c_vp * n)(*[vec(v, G) for v in g_weights])
    if Q and G:
        ctx.check(ctx.lib.dali_fuse_f32(ctx.h, dptr, n, wq, wg, c_vp(optr), Q, G, max(ld, 1)))
    return out


def topk_identify(distmat, k=20, largest=False, col_ids=None):
    """Per-row top-k of a ``[Q,G]`` matrix: ``(values [Q,k] fp32, ids [Q,k] int32)``.

    Order == ``torch.argsort(distmat, dim=1, stable=True)[:, :k]`` (validateModels.py:93);
    ``largest=True`` mirrors ``torch.topk(S, k, largest=True)`` (validateModels.py:180)."""
    d = as_matrix(distmat, np.float32, "distmat")
    Q, G = d.shape
    ctx = _ctx_for(d)
    ids = None
    if col_ids is not None:
        ids = as_matrix(col_ids, np.int32, "col_ids")
        if ids.shape != d.shape or ids.ld != d.ld or ids.device != d.device:
            raise ValueError("col_ids must match distmat in shape, layout and device")
    vals, vptr = _alloc_out((Q, k), d.device)
    idx, iptr = _alloc_out((Q, k), d.device, "int32")
    if Q:
        ctx.check(ctx.lib.dali_topk_f32(ctx.h, c_vp(d.ptr), Q, G, d.ld, int(k), 1 if largest else 0,
                                        c_vp(ids.ptr) if ids is not None else None, c_vp(vptr),
                                        c_vp(iptr)))
    return vals, idx


def topk_features(qf, gf, k=20, metric="cosine", precision=DEFAULT_PRECISION, normalize=None,
                  largest=False, g_base=0):
    """Fused distance + top-k for 1:N identification: never returns the ``[Q,G]`` matrix."""
    a = as_matrix(qf, np.float32, "qf")
    b = as_matrix(gf, np.float32, "gf")
    if a.shape[1] != b.shape[1] or a.ld != a.shape[1] or b.ld != b.shape[1]:
        raise ValueError("feature matrices must be contiguous with equal dimension")
    m = _enum(METRICS, metric, "metric")
    if normalize is None:
        normalize = m == METRICS["cosine"]
    ctx = _ctx_for(a, b)
    dev = _device_of(a, b)
    Q = a.shape[0]
    vals, vptr = _alloc_out((Q, k), dev)
    idx, iptr = _alloc_out((Q, k), dev, "int32")
    if Q:
        ctx.check(ctx.lib.dali_topk_features_f32(
            ctx.h, c_vp(a.ptr), Q, c_vp(b.ptr), b.shape[0], a.shape[1], m,
            _precision(precision, normalize, a.shape[1]), 1 if normalize else 0, int(k),
            1 if largest else 0, int(g_base), c_vp(vptr), c_vp(iptr)))
    return vals, idx


def re_ranking(q_g_dist, q_q_dist, g_g_dist, k1=20, k2=6, lambda_value=0.3):
    """k-reciprocal re-ranking with ``torchreid.utils.re_ranking``'s signature: the hook the
    reference keeps commented out after every distance matrix (validateModels.py:49-53,
    evaluate.py:294-298, evaluate_ensembled_models.py:284-288,303-307).

    ``q_g_dist`` [Q,G], ``q_q_dist`` [Q,Q], ``g_g_dist`` [G,G]; returns the re-ranked ``[Q,G]``
    float32 matrix where the inputs live (numpy in, numpy out; CUDA tensor in, CUDA tensor out)."""
    a = as_matrix(q_g_dist, np.float32, "q_g_dist")
    b = as_matrix(q_q_dist, np.float32, "q_q_dist")
    c = as_matrix(g_g_dist, np.float32, "g_g_dist")
    Q, G = a.shape
    if b.shape != (Q, Q) or c.shape != (G, G):
        raise ValueError("q_q_dist must be [Q,Q] and g_g_dist [G,G]")
    if len({a.device, b.device, c.device}) != 1:
        raise ValueError("the three matrices must live on the same device (or all on the host)")
    ctx = _ctx_for(a, b, c)
    out, optr = _alloc_out((Q, G), a.device)
    if Q and G:
        ctx.check(ctx.lib.dali_rerank_f32(ctx.h, c_vp(a.ptr), max(a.ld, 1), c_vp(b.ptr), max(b.ld, 1),
                                          c_vp(c.ptr), max(c.ld, 1), Q, G, int(k1), int(k2),
                                          float(lambda_value), c_vp(optr), max(G, 1)))
    return out


def evaluate_features(qf, gf, q_pids, g_pids, q_camids, g_camids, metric="cosine",
                      precision=DEFAULT_PRECISION, normalize=None, max_rank=50, accum="cy_f32",
                      return_distmat=False, return_details=False):
    """Features in, ``(cmc, mAP)`` out: normalise -> distance matrix -> rank -> CMC/mAP in one
    library call (the arithmetic of validateModels.validate, validateModels.py:41-47,61-69)."""
    a = as_matrix(qf, np.float32, "qf")
    b = as_matrix(gf, np.float32, "gf")
    if a.shape[1] != b.shape[1] or a.ld != a.shape[1] or b.ld != b.shape[1]:
        raise ValueError("feature matrices must be contiguous with equal dimension")
    Q, D = a.shape
    G = b.shape[0]
    qp, gp = canonicalize_labels(q_pids, g_pids)
    qc, gc = canonicalize_labels(q_camids, g_camids)
    if qp.shape[0] != Q or gp.shape[0] != G or qc.shape[0] != Q or gc.shape[0] != G:
        raise ValueError("label arrays do not match the feature matrices")
    if G < max_rank:
        max_rank = G
        print("Note: number of gallery samples is quite small, got {}".format(G))
    if Q == 0 or G == 0:
        raise AssertionError(_lib.NO_VALID_MSG)
    m = _enum(METRICS, metric, "metric")
    if normalize is None:
        normalize = m == METRICS["cosine"]
    ctx = _ctx_for(a, b)
    dev = _device_of(a, b)
    cmc = np.zeros(max_rank, dtype=np.float32)
    mAP = ctypes.c_double(0.0)
    ap = np.zeros(Q, dtype=np.float64)
    first = np.zeros(Q, dtype=np.int32)
    nvalid = ctypes.c_int64(0)
    dist, dptr = (None, None)
    if return_distmat:
        dist, dptr = _alloc_out((Q, G), dev)
    rc = ctx.lib.dali_eval_features_f32(
        ctx.h, a.ptr, Q, b.ptr, G, D, p_i32(qp), p_i32(gp), p_i32(qc), p_i32(gc), m,
        _precision(precision, normalize, D), 1 if normalize else 0, int(max_rank),
        _enum(ACCUMS, accum, "accumulation mode"), np_ptr(cmc), ctypes.byref(mAP), np_ptr(ap), np_ptr(first),
        ctypes.byref(nvalid), dptr, G)
    ctx.check(rc)
    res = (cmc, float(mAP.value))
    if return_distmat:
        res = res + (dist,)
    if return_details:
        res = res + ({"ap": ap, "first_rank": first, "num_valid": int(nvalid.value)},)
    return res


def mrfuse(score_mats, topk=20, use_columns=False, killscale=1.0, return_details=False):
    """Meta-recognition fusion of up to three ``[Q,G]`` SIMILARITY matrices (fp32):
    ``sum_m(w_m*s_m) / sum_m(w_m)`` with ``w_m`` the CDF of a Weibull fitted per gallery column --
    ``Meta_Recognition.mrfuse`` / ``metarec`` of evaluate.py:583-627.  Returns fp64 ``[Q,G]`` on the
    side the inputs live on; with ``return_details`` also ``{"weights": [n,Q,G] fp64,
    "fit": [n,G,2] (shape, scale), "small": [n,G]}``."""
    bufs = [as_matrix(m, np.float32, "scores") for m in score_mats]
    n = len(bufs)
    if not 1 <= n <= 3:
        raise ValueError("mrfuse takes one to three score matrices")
    Q, G = bufs[0].shape
    dev = bufs[0].device
    for b in bufs:
        if b.shape != (Q, G) or b.device != dev:
            raise ValueError("score matrices must have one shape and live on one device")
    if len({b.ld for b in bufs}) != 1:  # mixed padded / contiguous layouts: one leading dimension for all
        bufs = [as_matrix(b.keep.contiguous() if hasattr(b.keep, "contiguous") else np.ascontiguousarray(b.keep),
                          np.float32, "scores") for b in bufs]
    ctx = _ctx_for(*bufs)
    out, optr = _alloc_out((Q, G), dev, "float64")
    fit = small = weights = None
    fptr = sptr = wptr = None
    if return_details:
        weights, wptr = _alloc_out((n, Q, G), dev, "float64")
        fit = np.empty((n, G, 2), dtype=np.float64)
        small = np.empty((n, G), dtype=np.float32)
        fptr, sptr = fit.ctypes.data, small.ctypes.data
    ptrs = (c_vp * n)(*[c_vp(b.ptr) for b in bufs])
    ctx.check(ctx.lib.dali_mrfuse_f32(ctx.h, ptrs, n, Q, G, bufs[0].ld, int(topk), 1 if use_columns else 0,
                                      float(killscale), c_vp(optr), G, c_vp(fptr), c_vp(sptr), c_vp(wptr)))
    if return_details:
        return out, {"weights": weights, "fit": fit, "small": small}
    return out


def argsort_rows(distmat, descending=False):
    """``torch.argsort(distmat, dim=1, descending=descending, stable=True)`` as int32 ``[Q,G]``:
    the whole ranked list of every row (getFeatures.py:303,347; ``np.argsort(distmat, axis=1)`` of
    the torchreid evaluation).  Ties by ascending column; NaN after +inf.  Output lives where the
    input does."""
    d = as_matrix(distmat, np.float32, "distmat")
    Q, G = d.shape
    ctx = _ctx_for(d)
    idx, iptr = _alloc_out((Q, G), d.device, "int32")
    if Q and G:
        ctx.check(ctx.lib.dali_argsort_f32(ctx.h, c_vp(d.ptr), Q, G, d.ld, 1 if descending else 0, c_vp(iptr)))
    return idx

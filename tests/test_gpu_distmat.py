"""GPU parity tests of normalisation, the distance contraction (exact FP32 pipe, 3xTF32 and
single-pass TF32 on tcgen05), fusion and top-k through the C-ABI.

Tolerances (north_star): distances of the exact / fp32-class paths within 1e-5 relative of
the reference's CPU fp32 expression -- written as |delta| <= 1e-5 * max(1, |d|) because
1 - cos is ill-conditioned near 0; the TF32 path within 0.01 percentage points of mAP."""
import os

import numpy as np
import pytest
import torch

from oracle import distmat_oracle as do
from oracle import rank_oracle as ro

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-5


def _close(a, e, tol=TOL, scale=None):
    a = np.asarray(a.cpu() if isinstance(a, torch.Tensor) else a, dtype=np.float64)
    e = np.asarray(e, dtype=np.float64)
    den = np.maximum(1.0, np.abs(e)) if scale is None else np.maximum(scale, np.abs(e))
    err = np.abs(a - e) / den
    assert np.isfinite(a).all()
    assert err.max() <= tol, f"max err {err.max():.3e} > {tol}"


def _norm_scale(q, g, metric):
    """Un-normalised metrics: the same 1e-5 bound, relative to the magnitude of the operands
    (|q||g| for the dot product, |q|^2+|g|^2 for squared distances)."""
    qn = np.linalg.norm(np.asarray(q, dtype=np.float64), axis=1)[:, None]
    gn = np.linalg.norm(np.asarray(g, dtype=np.float64), axis=1)[None, :]
    if metric == "dot":
        return np.maximum(1.0, qn * gn)
    if metric == "sqeuclidean":
        return np.maximum(1.0, qn ** 2 + gn ** 2)
    if metric == "euclidean":
        return np.maximum(1.0, qn + gn)
    return None


def test_normalize_matches_reference():
    from daliid_b200 import metrics
    z = np.load(os.path.join(GOLDEN, "tiny.npz"))
    out, norms = metrics.normalize(z["qf"], return_norms=True)
    _close(out, z["qn"], 2e-7)
    _close(norms, z["q_norm"], 1e-6)
    out_d = metrics.normalize(torch.from_numpy(z["qf"]).cuda())
    assert out_d.is_cuda and np.array_equal(out_d.cpu().numpy(), out)
    zero = np.zeros((2, 8), dtype=np.float32)
    zero[1, 3] = 2.0
    o = metrics.normalize(zero)
    assert np.isnan(o[0]).all() and o[1, 3] == 1.0  # no eps, like the reference


@pytest.mark.parametrize("precision", ["fp32", "tf32x3", "tf32c"])
@pytest.mark.parametrize("metric", ["cosine", "sqeuclidean", "euclidean", "dot"])
def test_golden_distances(metric, precision):
    from daliid_b200 import metrics
    z = np.load(os.path.join(GOLDEN, "tiny.npz"))
    out = metrics.compute_distance_matrix(z["qf"], z["gf"], metric=metric, precision=precision)
    assert out.dtype == np.float32 and out.shape == z[metric].shape
    _close(out, z[metric], scale=_norm_scale(z["qf"], z["gf"], metric))
    out_d = metrics.compute_distance_matrix(torch.from_numpy(z["qf"]).cuda(),
                                            torch.from_numpy(z["gf"]).cuda(), metric, precision)
    assert out_d.is_cuda and np.array_equal(out_d.cpu().numpy(), out)


@pytest.mark.parametrize("Q,G,D", [(1, 1, 1), (5, 3, 7), (130, 257, 33), (128, 256, 32),
                                   (300, 1000, 768), (129, 513, 2048), (257, 300, 3840)])
@pytest.mark.parametrize("precision", ["fp32", "tf32x3", "tf32c", "f16x3"])
def test_random_shapes_cosine(Q, G, D, precision):
    from daliid_b200 import metrics
    g = torch.Generator().manual_seed(Q * 7 + G)
    qf = torch.randn(Q, D, generator=g)
    gf = torch.randn(G, D, generator=g) + 0.3
    ref = do.cosine_distmat(qf, gf).numpy()
    out = metrics.compute_distance_matrix(qf.cuda(), gf.cuda(), "cosine", precision)
    _close(out, ref)


def test_f16x3_golden_and_stress():
    """fp16 hi/lo split (rows scaled by 2^12): golden cosine matrix, unit-row squared Euclidean,
    badly scaled rows (1e-6 .. 1e6, one dominant coordinate, many tiny ones), and the refusal of
    un-normalised operands."""
    from daliid_b200 import metrics, _lib
    z = np.load(os.path.join(GOLDEN, "tiny.npz"))
    out = metrics.compute_distance_matrix(z["qf"], z["gf"], "cosine", "f16x3")
    _close(out, z["cosine"])
    qn = torch.from_numpy(z["qn"])
    gn = torch.from_numpy(z["gf"])
    gn = gn / torch.norm(gn, dim=1, keepdim=True)
    ref = (qn ** 2).sum(1)[:, None] + (gn ** 2).sum(1)[None, :] - 2.0 * qn @ gn.T
    out = metrics.compute_distance_matrix(z["qf"], z["gf"], "sqeuclidean", "f16x3", normalize=True)
    _close(out, ref.numpy())
    g = torch.Generator().manual_seed(5)
    qf = torch.randn(257, 2048, generator=g)
    gf = torch.randn(700, 2048, generator=g)
    qf[:64] *= 1e6
    qf[64:128] *= 1e-6
    gf[:100] *= 3e5
    gf[100:200, 0] = 50.0          # one dominant coordinate, the rest ~1/50 of it
    gf[200:300] *= torch.logspace(-6, 0, 2048)[None, :]   # six decades inside a row
    ref = do.cosine_distmat(qf, gf).numpy()
    out = metrics.compute_distance_matrix(qf.cuda(), gf.cuda(), "cosine", "f16x3")
    _close(out, ref)
    with pytest.raises(_lib.DaliError):
        metrics.compute_distance_matrix(z["qf"], z["gf"], "sqeuclidean", "f16x3")


@pytest.mark.parametrize("D", [512, 768, 1280, 2048, 3840, 4096])
def test_default_precision_duplicates_within_1e5(D):
    """The parity bar (distances within 1e-5 of the reference's fp32 expression) on the pairs where
    the tensor core's accumulator truncation shows: exact duplicates, near-duplicates, and
    adversarial partial-sum trajectories (nearly all of |x|^2 in the first / in the last 32
    coordinates), at every feature size of the reference.  No exemption at any D (DESIGN.md 4.1:
    fixed-point hi plane, compensation of the mean loss, corrections-first schedule above 2048)."""
    from daliid_b200 import metrics
    g = torch.Generator().manual_seed(D)
    a = torch.randn(192, D, generator=g)
    head = a[:32].clone()
    head[:, 32:] *= 0.02
    tail = a[32:64].clone()
    tail[:, :-32] *= 0.02
    q = torch.cat([a, head, tail])
    gal = torch.cat([a[:96], a[:96] + 0.01 * torch.randn(96, D, generator=g), head, tail,
                     torch.randn(200, D, generator=g)])
    ref = do.cosine_distmat(q, gal).numpy()
    ref64 = (1.0 - (q.double() / q.double().norm(dim=1, keepdim=True)) @
             (gal.double() / gal.double().norm(dim=1, keepdim=True)).T).numpy()
    for prec in ("auto", "f16x3"):
        out = metrics.compute_distance_matrix(q.cuda(), gal.cuda(), "cosine", prec).cpu().numpy()
        _close(out, ref, 1e-5)
        # against float64 the margin is visible: 8.5e-6 even for the adversarial rows
        assert np.abs(out - ref64).max() <= 8.5e-6, np.abs(out - ref64).max()


def test_tf32c_duplicates_documented_bound():
    """TF32C is the non-default mode for un-normalised operands; its accumulator loss on exact
    duplicates is D/4 * 2^-24 -- within 1e-5 up to D = 768, stated (not hidden)
    above.  Ordinary pairs are within 1e-5 at every D (test_random_shapes_cosine)."""
    from daliid_b200 import metrics
    g = torch.Generator().manual_seed(5)
    for D, tol in ((768, 1e-5), (2048, 2048 / 4 * 2.0 ** -24)):
        a = torch.randn(64, D, generator=g)
        b = torch.cat([a[:32], torch.randn(100, D, generator=g)])
        ref = do.cosine_distmat(a, b).numpy()
        out = metrics.compute_distance_matrix(a.cuda(), b.cuda(), "cosine", "tf32c")
        _close(out, ref, tol)


def test_auto_precision_routes_long_rows_to_fp32():
    """Unit rows longer than 4096 elements leave the range where f16x3 is within 1e-5 for
    duplicates: "auto" takes the exact FP32 pipe there."""
    from daliid_b200 import metrics
    assert metrics._precision("auto", True, 4096) == metrics.PRECISIONS["f16x3"]
    assert metrics._precision("auto", True, 4097) == metrics.PRECISIONS["fp32"]
    assert metrics._precision("auto", False, 8192) == metrics.PRECISIONS["tf32c"]
    g = torch.Generator().manual_seed(1)
    a = torch.randn(40, 5000, generator=g)
    ref = do.cosine_distmat(a, a).numpy()
    out = metrics.compute_distance_matrix(a.cuda(), a.cuda(), "cosine").cpu().numpy()
    _close(out, ref, 1e-5)


def test_tf32_single_pass_quality():
    """Single-pass TF32 on a small case: distances within 2e-3 absolute and mAP within 0.05 pp
    of the exact path (the 0.01 pp claim is checked at the Market shape in
    test_gpu_full_size.py, where 3368 queries average the rank noise)."""
    from daliid_b200 import metrics, synth
    qf, gf, qp, gp, qc, gc = synth.make_config("small", device="cuda")
    exact = metrics.compute_distance_matrix(qf, gf, "cosine", "fp32")
    m_exact = metrics.evaluate_rank(exact, qp, gp, qc, gc)[1]
    for prec in ("tf32", "f16"):  # both keep 11 mantissa bits of the unit rows
        fast = metrics.compute_distance_matrix(qf, gf, "cosine", prec)
        assert (exact - fast).abs().max().item() < 2e-3
        m_fast = metrics.evaluate_rank(fast, qp, gp, qc, gc)[1]
        assert abs(m_exact - m_fast) * 100 <= 0.05


def test_exact_path_is_tile_position_independent():
    """A gallery slab computed alone equals the same columns of the full matrix bit for bit
    (what makes gallery sharding exact)."""
    from daliid_b200 import metrics
    g = torch.Generator().manual_seed(3)
    qf = torch.randn(200, 384, generator=g).cuda()
    gf = torch.randn(1500, 384, generator=g).cuda()
    for precision in ("fp32", "tf32x3", "tf32c", "tf32", "f16x3", "f16"):
        full = metrics.compute_distance_matrix(qf, gf, "cosine", precision)
        part = metrics.compute_distance_matrix(qf, gf[700:1333].contiguous(), "cosine", precision)
        assert torch.equal(full[:, 700:1333], part), precision


def test_evaluate_features_end_to_end():
    """Fused call: bit-exact CMC/mAP w.r.t. the oracle fed the SAME distance matrix, and within
    rounding of the oracle fed the reference's CPU matrix."""
    from daliid_b200 import metrics, synth
    qf, gf, qp, gp, qc, gc = synth.make_config("small")
    for precision in ("fp32", "tf32x3", "tf32c", "f16x3", "auto"):
        cmc, mAP, dist, det = metrics.evaluate_features(qf, gf, qp, gp, qc, gc, precision=precision,
                                                        return_distmat=True, return_details=True)
        e = ro.eval_market1501_cy_f32(dist, qp, gp, qc, gc, return_details=True)
        assert np.array_equal(cmc, e[0]) and mAP == e[1]
        assert np.array_equal(det["first_rank"], e[3])
        ref = ro.eval_market1501_cy_f32(do.cosine_distmat(qf, gf).numpy(), qp, gp, qc, gc)
        assert abs(ref[1] - mAP) * 100 <= 0.01 and np.abs(ref[0] - cmc).max() <= 0.005
        _close(dist, do.cosine_distmat(qf, gf).numpy())
    # device-resident features give the same bits as host features (same arithmetic on both sides)
    for precision in ("tf32x3", "auto"):
        c1, m1 = metrics.evaluate_features(qf, gf, qp, gp, qc, gc, precision=precision)
        c2, m2 = metrics.evaluate_features(qf.cuda(), gf.cuda(), qp, gp, qc, gc, precision=precision)
        assert np.array_equal(c2, c1) and m2 == m1


def test_fusion_bit_exact():
    from daliid_b200 import metrics
    z = np.load(os.path.join(GOLDEN, "fusion.npz"))
    for dev in ("host", "cuda"):
        ds = [z["d0"], z["d1"], z["d2"]]
        if dev == "cuda":
            ds = [torch.from_numpy(d).cuda() for d in ds]
        get = lambda t: t.cpu().numpy() if isinstance(t, torch.Tensor) else t
        assert np.array_equal(get(metrics.fuse_distmats(ds[:2])), z["mean2"])
        assert np.array_equal(get(metrics.fuse_distmats(ds)), z["mean3"])
        w = metrics.fuse_distmats(ds[:2], q_weights=[z["qm0"], z["qm1"]], g_weights=[z["gm0"], z["gm1"]])
        assert np.array_equal(get(w), z["weighted"])


@pytest.mark.parametrize("shape", [(37, 95, 64), (300, 1000, 768), (513, 777, 96)])
@pytest.mark.parametrize("nmodels", [1, 2, 3, 5])
@pytest.mark.parametrize("precision", ["f16x3", "tf32", "tf32c"])
def test_ensemble_mean_in_the_epilogue_is_bit_identical(shape, nmodels, precision):
    """evaluate.py:260-278 / evaluate_ensembled_models.py:313: the mean formed in the contractions'
    epilogues equals fuse_distmats of the separately materialised matrices bit for bit, and the
    individual matrices equal compute_distance_matrix's."""
    from daliid_b200 import metrics
    Q, G, D = shape
    rng = np.random.default_rng(Q + nmodels)
    qs = [torch.from_numpy(rng.standard_normal((Q, D)).astype(np.float32)).cuda() for _ in range(nmodels)]
    gs = [torch.from_numpy(rng.standard_normal((G, D)).astype(np.float32)).cuda() for _ in range(nmodels)]
    sep = [metrics.compute_distance_matrix(q, g, "cosine", precision) for q, g in zip(qs, gs)]
    ref = metrics.fuse_distmats(sep)
    ds, mean = metrics.ensemble_distance_matrices(qs, gs, "cosine", precision)
    assert mean.shape == (Q, G) and len(ds) == nmodels
    assert torch.equal(mean, ref)
    for a, b in zip(ds, sep):
        assert torch.equal(a, b)
    none, mean_only = metrics.ensemble_distance_matrices(qs, gs, "cosine", precision, individual=False)
    assert none is None and torch.equal(mean_only, ref)
    # the reference's own expression on the CPU (the matrices themselves are the contraction's)
    cpu = sep[0].cpu().numpy().copy()
    for d in sep[1:]:
        cpu = cpu + d.cpu().numpy()
    assert np.array_equal(mean.cpu().numpy(), cpu / np.float32(nmodels))


@pytest.mark.parametrize("n", [2, 3, 5, 6, 7, 8])
def test_mean_division_is_the_ieee_division_for_every_operand(n):
    """The mean's x / n (fuse.cu, fused-mean epilogue) is a reciprocal multiplication refined by two
    FMAs; exhaustively equal to the IEEE division numpy and torch perform (all 2^32 operands)."""
    import ctypes
    from daliid_b200 import _lib
    ctx = _lib.get_ctx(0)
    bad = ctypes.c_uint64(123)
    ctx.check(ctx.lib.dali_selftest_mean_division(ctx.h, n, ctypes.byref(bad)))
    assert bad.value == 0


def test_ensemble_falls_back_for_the_fp32_pipe_and_host_results():
    from daliid_b200 import metrics
    rng = np.random.default_rng(3)
    qs = [rng.standard_normal((20, 48)).astype(np.float32) for _ in range(2)]
    gs = [rng.standard_normal((33, 48)).astype(np.float32) for _ in range(2)]
    ds, mean = metrics.ensemble_distance_matrices(qs, gs, "cosine", "fp32")
    assert isinstance(mean, np.ndarray) and np.array_equal(mean, metrics.fuse_distmats(ds))
    qc, gc = [torch.from_numpy(x).cuda() for x in qs], [torch.from_numpy(x).cuda() for x in gs]
    ds, mean = metrics.ensemble_distance_matrices(qc, gc, "cosine", "fp32")
    assert torch.equal(mean, metrics.fuse_distmats(ds))
    ds, mean = metrics.ensemble_distance_matrices(qc, gc, "sqeuclidean", "tf32c")
    sep = [metrics.compute_distance_matrix(q, g, "sqeuclidean", "tf32c") for q, g in zip(qc, gc)]
    assert torch.equal(mean, metrics.fuse_distmats(sep))


def test_fusion_unaligned_shape():
    from daliid_b200 import metrics
    rng = np.random.default_rng(1)
    ds = [rng.random((7, 13)).astype(np.float32) for _ in range(3)]
    assert np.array_equal(metrics.fuse_distmats(ds), do.fuse_mean(ds))


@pytest.mark.parametrize("k", [1, 5, 20, 128])
def test_topk_matches_stable_argsort(k):
    from daliid_b200 import metrics
    z = np.load(os.path.join(GOLDEN, "ties.npz"))
    d = z["dist"]
    vals, idx = metrics.topk_identify(d, k=k)
    order = ro.stable_argsort(d)[:, :k]
    assert np.array_equal(idx, order.astype(np.int32))
    assert np.array_equal(vals, np.take_along_axis(d, order, 1), equal_nan=True)
    if k == 20:
        assert np.array_equal(idx, z["top20"])
    rng = np.random.default_rng(k)
    big = rng.random((37, 20011)).astype(np.float32)
    big = (np.round(big * 4096) / 4096).astype(np.float32)
    v2, i2 = metrics.topk_identify(torch.from_numpy(big).cuda(), k=k)
    assert np.array_equal(i2.cpu().numpy(), ro.stable_argsort(big)[:, :k].astype(np.int32))


@pytest.mark.parametrize("largest", [False, True])
@pytest.mark.parametrize("G", [256, 257, 1003, 4099, 15913, 16378])
def test_topk_rows_in_registers(G, largest):
    """``topk_minima_kernel`` (rows of 256 .. 16378 columns held in registers, threshold from group
    minima): every alignment of the row start (odd leading dimension), rows with NaN / +-inf / -0,
    quantised values with many ties at the threshold, constant rows (more candidates than the list
    holds: the repair kernel), rows with fewer than k numbers -- against the stable argsort order
    (smallest) and torch.topk's order (largest: NaN first, ties by ascending index)."""
    from daliid_b200 import metrics
    rng = np.random.default_rng(G + int(largest))
    Q = 11
    d = rng.standard_normal((Q, G)).astype(np.float32)
    d[1] = np.round(d[1] * 8) / 8                      # heavy ties
    d[2] = 0.25                                        # constant row
    d[3, ::3] = np.nan
    d[4] = np.nan
    d[4, 5:9] = [3.0, -0.0, 0.0, -np.inf]              # fewer numbers than k
    d[5, :7] = [np.inf, -np.inf, -0.0, 0.0, np.nan, 1e-38, -1e-38]
    d[6] = np.sort(d[6])                               # best first
    d[7] = np.sort(d[7])[::-1]                         # best last
    d[8, G // 2:] = d[8, : G - G // 2]                 # duplicated halves
    ld = G + 3
    buf = torch.full((Q * ld + 4,), 7.0, dtype=torch.float32, device="cuda")
    for shift in range(4):                             # rows start at every 4-byte phase of 16 bytes
        view = buf[shift: shift + Q * ld].view(Q, ld)[:, :G]
        view.copy_(torch.from_numpy(d))
        for k in (1, 20, 64, 128):
            v, i = metrics.topk_identify(view, k=k, largest=largest)
            if largest:  # torch.topk: NaN above +inf, then descending, ties by ascending index
                nan = np.isnan(d)
                val = np.where(nan, 0.0, d) + 0.0
                order = np.stack([np.lexsort((np.arange(G), -val[r], ~nan[r])) for r in range(Q)])[:, :k]
            else:
                order = ro.stable_argsort(d)[:, :k]
            assert np.array_equal(i.cpu().numpy(), order.astype(np.int32)), (G, largest, shift, k)
            assert np.array_equal(v.cpu().numpy(), np.take_along_axis(d, order, 1), equal_nan=True)


@pytest.mark.parametrize("largest", [False, True])
@pytest.mark.parametrize("G", [30011, 40013])
def test_topk_streaming_long_rows(largest, G):
    """Rows of >= 4096 columns take the streaming filter + compaction path; rows whose candidate
    list overflows (here: values quantised to 9 levels, one row sorted so that every later column
    is better, a row of NaN / inf) are redone on the device by the one-CTA-per-row kernel.  Equal
    to the stable argsort prefix in all cases, odd leading dimension included."""
    from daliid_b200 import metrics
    rng = np.random.default_rng(21)
    Q = 41  # (30011 columns: the row-in-registers kernel with 1024 threads; 40013: the streaming path)
    d = rng.random((Q, G)).astype(np.float32)
    d[1] = np.round(d[1] * 8) / 8                       # massive ties
    d[2] = np.sort(d[2])[::-1] if not largest else np.sort(d[2])   # every later column is better
    d[3] = np.nan
    d[4, ::3] = np.inf
    d[5, ::5] = -np.inf
    d[6, 100:200] = np.nan
    d[7] = 0.25                                          # one value only
    sign = -1.0 if largest else 1.0
    dt = torch.from_numpy(d).cuda()
    for k in (1, 20, 128):
        v, i = metrics.topk_identify(dt, k=k, largest=largest)
        if largest:
            # torch.topk semantics: NaN counts as the largest value, ties by ascending index
            key = np.where(np.isnan(d), np.inf, d)
            order = np.argsort(-key, axis=1, kind="stable")[:, :k]
        else:
            order = ro.stable_argsort(d)[:, :k]
        assert np.array_equal(i.cpu().numpy(), order.astype(np.int32)), (largest, k)
        assert np.array_equal(v.cpu().numpy(), np.take_along_axis(d, order, 1), equal_nan=True)
    # host matrix with an odd width (unaligned rows)
    v, i = metrics.topk_identify(d[:, :G - 10].copy(), k=20, largest=largest)
    e, ei = metrics.topk_identify(dt[:, :G - 10].contiguous(), k=20, largest=largest)
    assert np.array_equal(i, ei.cpu().numpy()) and np.array_equal(v, e.cpu().numpy(), equal_nan=True)


def test_topk_largest_short_rows_and_merge():
    from daliid_b200 import metrics
    rng = np.random.default_rng(9)
    s = rng.random((11, 300)).astype(np.float32)
    v, i = metrics.topk_identify(s, k=5, largest=True)
    e = torch.topk(torch.from_numpy(s), 5, dim=1, largest=True)
    assert np.array_equal(i, e.indices.numpy().astype(np.int32)) and np.array_equal(v, e.values.numpy())
    v, i = metrics.topk_identify(s[:, :3], k=5)  # G < k: padded with (+inf, -1)
    assert (i[:, 3:] == -1).all() and np.isinf(v[:, 3:]).all()
    assert np.array_equal(i[:, :3], ro.stable_argsort(s[:, :3]).astype(np.int32))
    # k-way merge of per-slab candidates through col_ids == global top-k
    parts = [(0, 100), (100, 100), (200, 100)]
    cv, ci = [], []
    for g0, gs in parts:
        pv, pi = metrics.topk_identify(s[:, g0:g0 + gs], k=7)
        cv.append(pv); ci.append(pi + g0)
    mv, mi = metrics.topk_identify(np.concatenate(cv, 1), k=7, col_ids=np.concatenate(ci, 1))
    assert np.array_equal(mi, ro.stable_argsort(s)[:, :7].astype(np.int32))


def test_topk_features_fused():
    from daliid_b200 import metrics, synth
    qf, gf, qp, gp, qc, gc = synth.make_config("small", device="cuda")
    d = metrics.compute_distance_matrix(qf, gf, "cosine", "tf32x3")
    v, i = metrics.topk_features(qf, gf, k=20, precision="tf32x3", g_base=1000)
    ev, ei = metrics.topk_identify(d, k=20)
    assert torch.equal(v, ev) and torch.equal(i, ei + 1000)


@pytest.mark.parametrize("precision,metric,largest", [
    ("f16x3", "cosine", False), ("tf32c", "cosine", False), ("tf32x3", "cosine", False),
    ("tf32", "cosine", False), ("tf32c", "sqeuclidean", False), ("tf32", "euclidean", False),
    ("tf32c", "dot", True), ("tf32", "dot", True), ("f16x3", "cosine", True), ("f16", "cosine", False)])
def test_topk_features_fused_multi_chunk(precision, metric, largest):
    """Fused distance + top-k (several gallery chunks, ragged last tile) equals top-k of the
    materialised matrix of the same precision: same kernel arithmetic, so bit-identical."""
    from daliid_b200 import metrics
    g = torch.Generator().manual_seed(12)
    Q, G, D = 300, 7777, 96
    qf = torch.randn(Q, D, generator=g).cuda()
    gf = torch.randn(G, D, generator=g).cuda()
    gf[100] = gf[5000]                      # exact duplicates: ties broken by gallery index
    gf[7000] = gf[5000]
    d = metrics.compute_distance_matrix(qf, gf, metric, precision)
    from daliid_b200 import _lib
    n_fb = _lib.get_ctx(0).fallback_count()
    for k in (1, 20, 128):
        v, i = metrics.topk_features(qf, gf, k=k, metric=metric, precision=precision,
                                     largest=largest, g_base=7)
        ev, ei = metrics.topk_identify(d, k=k, largest=largest)
        assert torch.equal(i, ei + 7), (precision, metric, k)
        assert torch.equal(v, ev)
    assert _lib.get_ctx(0).fallback_count() == n_fb  # the fused path itself produced these


def test_topk_features_fused_overflow_falls_back():
    """Gallery ordered so that every later item is closer to every query than all earlier ones:
    each chunk overflows the candidate lists and the call must redo itself unfused."""
    from daliid_b200 import metrics
    g = torch.Generator().manual_seed(3)
    Q, G, D = 64, 6000, 64
    u = torch.nn.functional.normalize(torch.randn(1, D, generator=g), dim=1)
    v = torch.nn.functional.normalize(torch.randn(G, D, generator=g), dim=1)
    v = torch.nn.functional.normalize(v - (v @ u.T) * u, dim=1)
    theta = torch.linspace(1.5, 0.05, G)[:, None]
    gf = (torch.cos(theta) * u + torch.sin(theta) * v).cuda()
    qf = (u + 1e-3 * torch.randn(Q, D, generator=g)).cuda()
    d = metrics.compute_distance_matrix(qf, gf, "cosine", "tf32c")
    from daliid_b200 import _lib
    n_fb = _lib.get_ctx(0).fallback_count()
    vv, ii = metrics.topk_features(qf, gf, k=20, precision="tf32c")
    assert _lib.get_ctx(0).fallback_count() == n_fb + 1
    ev, ei = metrics.topk_identify(d, k=20)
    assert torch.equal(ii, ei) and torch.equal(vv, ev)
    assert int(ii.min()) > G - 200          # the best items really are the last ones


def test_two_cta_kernel_equals_one_cta_kernel():
    """The CTA-pair contraction (plain and L2-banded tile order; TMA-store and transposing
    epilogue) and the one-CTA-per-tile kernel accumulate every element in the same k order:
    bit-identical matrices (subprocesses with DALI_UMMA_2CTA=0 / DALI_UMMA_NBAND=3 /
    DALI_UMMA_TMA_STORE=0).  The 'small' gallery has 2100 rows: a multiple of 4, so the default
    run stores through TMA."""
    import subprocess
    import sys
    import tempfile
    code = (
        "import torch, sys, numpy as np\n"
        "from daliid_b200 import metrics, synth\n"
        "qf, gf, *_ = synth.make_config('small', device='cuda')\n"
        "qf = torch.cat([qf, qf * 0.5 + 1.0, qf[:177] - 2.0])  # 777 queries: 4 pair tiles\n"
        "for p in sys.argv[2:]:\n"
        "    d = metrics.compute_distance_matrix(qf, gf, 'cosine', p)\n"
        "    np.save(sys.argv[1] + p + '.npy', d.cpu().numpy())\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    runs = {"one": (dict(DALI_UMMA_2CTA="0"), ("tf32", "tf32x3", "tf32c")),
            "two": (dict(), ("tf32", "tf32x3", "tf32c", "f16x3")),
            "band": (dict(DALI_UMMA_NBAND="3"), ("tf32", "tf32x3", "tf32c", "f16x3")),
            "notma": (dict(DALI_UMMA_TMA_STORE="0"), ("tf32", "tf32x3", "tf32c", "f16x3"))}
    with tempfile.TemporaryDirectory() as td:
        for name, (env, precs) in runs.items():
            subprocess.check_call([sys.executable, "-c", code, os.path.join(td, name + "_"), *precs],
                                  env=dict(os.environ, **env), cwd=root)
        for p in ("tf32", "tf32x3", "tf32c", "f16x3"):
            b = np.load(os.path.join(td, "two_" + p + ".npy"))
            assert np.array_equal(np.load(os.path.join(td, "band_" + p + ".npy")), b), p
            assert np.array_equal(np.load(os.path.join(td, "notma_" + p + ".npy")), b), p
            if p != "f16x3":
                assert np.array_equal(np.load(os.path.join(td, "one_" + p + ".npy")), b), p


def test_property_topk_and_fusion_hypothesis():
    """Property test (hypothesis, derandomised): top-k of arbitrary small matrices full of ties,
    signed zeros, infinities and NaN equals the stable argsort prefix (smallest) / torch.topk
    order (largest); mean fusion of 2-4 matrices equals numpy's left-to-right expression bitwise."""
    from hypothesis import given, settings, strategies as st, HealthCheck
    from daliid_b200 import metrics
    specials = np.array([0.0, -0.0, 0.5, 0.5, 1.0, np.inf, -np.inf, np.nan, -2.0, 3.0], dtype=np.float32)

    @settings(max_examples=80, deadline=None, derandomize=True,
              suppress_health_check=[HealthCheck.too_slow])
    @given(st.integers(1, 9), st.integers(1, 200), st.integers(1, 40), st.integers(0, 2 ** 31 - 1),
           st.booleans(), st.integers(2, 4))
    def run(Q, G, k, seed, largest, nmat):
        rng = np.random.default_rng(seed)
        d = rng.choice(specials, size=(Q, G)) if seed % 2 else rng.random((Q, G), dtype=np.float32)
        d = d.astype(np.float32)
        v, i = metrics.topk_identify(d, k=k, largest=largest)
        kk = min(k, G)
        if largest:  # torch.topk order: NaN above +inf, then descending, ties by ascending index
            order = np.array([sorted(range(G), key=lambda j: (0 if np.isnan(r[j]) else 1,
                                                              -(float(r[j]) + 0.0) if not np.isnan(r[j]) else 0.0, j))
                              for r in d])[:, :kk]
        else:
            order = ro.stable_argsort(d)[:, :kk]
        assert np.array_equal(i[:, :kk], order.astype(np.int32))
        assert np.array_equal(v[:, :kk], np.take_along_axis(d, order, 1), equal_nan=True)
        assert (i[:, kk:] == -1).all()
        ms = [rng.random((Q, G), dtype=np.float32) * 3 for _ in range(nmat)]
        acc = ms[0]
        for m in ms[1:]:
            acc = acc + m
        assert np.array_equal(metrics.fuse_distmats(ms), acc / nmat)

    run()

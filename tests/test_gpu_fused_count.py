"""Fused distance + positive-rank counting (the Q x G matrix is never written): bit-identical to the
matrix path of the same arithmetic, and -- through it -- to the oracle fed that matrix."""
import numpy as np
import pytest
import torch

from oracle import rank_oracle as ro

pytestmark = pytest.mark.gpu


def _fits(qp, gp, D):
    """The eligibility rule of the fused path (capi.cu fused_max_matches): at most 64 same-identity
    gallery items per query (128 from D = 1536 on)."""
    cnt = np.bincount(np.asarray(gp), minlength=1)
    qq = np.asarray(qp)
    qq = qq[qq < len(cnt)]
    max_m = int(cnt[qq].max()) if len(qq) else 0
    return 0 < max_m <= (128 if (D + 31) // 32 * 32 >= 1536 else 64)


def _small(device="cuda", **kw):
    """The "small" synthetic config with enough identities that a query has ~18 matches."""
    from daliid_b200 import synth
    cfg = dict(Q=300, G=2100, D=200, n_ids=120, n_cams=5, sigma=2.0)
    cfg.update(kw)
    return synth.make_features(seed=12, device=device, **cfg)


def _both(qf, gf, qp, gp, qc, gc, precision="f16x3", metric="cosine", accum="cy_f32", expect_fused=None):
    from daliid_b200 import _lib, metrics
    if expect_fused is None:
        expect_fused = _fits(qp, gp, qf.shape[1])
    ctx = _lib.get_ctx(qf.device.index if qf.is_cuda else 0)
    ctx.fused_count_enable(True)   # opt-in (the default is the matrix path)
    n0, f0 = ctx.fused_count_calls(), ctx.fallback_count()
    try:
        a = metrics.evaluate_features(qf, gf, qp, gp, qc, gc, metric=metric, precision=precision, accum=accum,
                                      return_details=True)
    finally:
        ctx.fused_count_enable(False)
    took = ctx.fused_count_calls() - n0
    assert took == (1 if expect_fused else 0), (took, ctx.fallback_count() - f0)
    b = metrics.evaluate_features(qf, gf, qp, gp, qc, gc, metric=metric, precision=precision, accum=accum,
                                  return_details=True, return_distmat=True)
    assert np.array_equal(a[2]["first_rank"], b[3]["first_rank"])
    assert np.array_equal(a[2]["ap"], b[3]["ap"], equal_nan=True)
    assert np.array_equal(a[0], b[0]) and a[1] == b[1]
    assert a[2]["num_valid"] == b[3]["num_valid"]
    return a, b


@pytest.mark.parametrize("name", ["tiny", "small"])
@pytest.mark.parametrize("precision", ["f16x3", "tf32c", "tf32", "f16"])
def test_fused_equals_matrix_path_and_oracle(name, precision):
    from daliid_b200 import synth
    qf, gf, qp, gp, qc, gc = synth.make_config(name, device="cuda") if name == "tiny" else _small()
    assert _fits(qp, gp, qf.shape[1])
    for accum in ("cy_f32", "py_f64"):
        a, b = _both(qf, gf, qp, gp, qc, gc, precision, accum=accum)
        d = b[2].cpu().numpy()
        fn = ro.eval_market1501_cy_f32 if accum == "cy_f32" else ro.eval_market1501_py_f64
        e = fn(d, qp, gp, qc, gc, 50, return_details=True)
        assert np.array_equal(a[0], e[0]) and a[1] == e[1] and np.array_equal(a[2]["first_rank"], e[3])


@pytest.mark.parametrize("metric", ["sqeuclidean", "euclidean"])
def test_fused_other_metrics(metric):
    qf, gf, qp, gp, qc, gc = _small()
    a, b = _both(qf, gf, qp, gp, qc, gc, "tf32c", metric=metric, expect_fused=True)


def test_fused_heavy_ties():
    """Duplicated gallery rows (exact distance ties between positives and negatives, and among
    positives) and quantised features: ranks are decided by gallery id, as in the matrix path."""
    qf, gf, qp, gp, qc, gc = _small()
    gf = gf.clone()
    g = torch.Generator().manual_seed(3)
    src = torch.randint(0, gf.shape[0], (600,), generator=g)
    dst = torch.randint(0, gf.shape[0], (600,), generator=g)
    gf[dst.cuda()] = gf[src.cuda()]                  # 600 duplicated rows, labels unchanged
    qf = (qf * 2).round() / 2                        # coarse features: many equal dot products
    gf = (gf * 2).round() / 2
    a, b = _both(qf, gf, qp, gp, qc, gc, "f16x3", expect_fused=True)
    d = b[2].cpu().numpy()
    assert (np.diff(np.sort(d, axis=1), axis=1) == 0).sum() > 1000   # the case really has ties
    e = ro.eval_market1501_cy_f32(d, qp, gp, qc, gc, 50, return_details=True)
    assert np.array_equal(a[2]["first_rank"], e[3]) and a[1] == e[1]


def test_fused_edge_cases():
    """Queries without any match / with only junk matches, a gallery narrower than a tile, a ragged
    last tile, more than 32 positives per query (second threshold pass at D >= 1536)."""
    from daliid_b200 import synth
    rng = np.random.default_rng(5)
    for Q, G, D, ids, cams in ((70, 200, 64, 15, 2), (300, 1000, 256, 30, 3), (513, 2049, 128, 300, 2),
                               (130, 900, 1536, 12, 4)):
        qf, gf, qp, gp, qc, gc = synth.make_features(Q, G, D, ids, cams, 2.0, seed=Q, device="cuda")
        qp = qp.copy()
        qp[::7] = 10_000 + np.arange(len(qp[::7]))   # identities absent from the gallery
        _both(qf, gf, qp, gp, qc, gc, "f16x3", expect_fused=True)
        if D in (256, 1536):
            assert np.bincount(gp).max() > 32        # further threshold passes are exercised


def test_fused_all_junk_raises_like_the_matrix_path():
    """One camera: every same-identity gallery item is junk, no query is valid -- torchreid's
    AssertionError on both paths."""
    from daliid_b200 import _lib, metrics, synth
    qf, gf, qp, gp, qc, gc = synth.make_features(100, 600, 64, 40, 1, 2.0, seed=3, device="cuda")
    ctx = _lib.get_ctx(0)
    for on in (True, False):
        ctx.fused_count_enable(on)
        try:
            with pytest.raises(AssertionError):
                metrics.evaluate_features(qf, gf, qp, gp, qc, gc, precision="f16x3")
        finally:
            ctx.fused_count_enable(False)


def test_fused_falls_back_on_non_finite_thresholds_and_many_matches():
    from daliid_b200 import _lib, synth
    qf, gf, qp, gp, qc, gc = _small()
    gf = gf.clone()
    gf[5] = 0.0                                      # zero-norm gallery row: NaN distances (reference: 0/0)
    ctx = _lib.get_ctx(0)
    f0 = ctx.fallback_count()
    _both(qf, gf, qp, gp, qc, gc, "f16x3", expect_fused=False)
    assert ctx.fallback_count() == f0 + 1
    # 200 gallery items per identity: beyond what the epilogue holds -> matrix path, no fallback
    qf, gf, qp, gp, qc, gc = synth.make_features(50, 2000, 200, 10, 3, 2.0, seed=1, device="cuda")
    f0 = ctx.fallback_count()
    _both(qf, gf, qp, gp, qc, gc, "f16x3", expect_fused=False)
    assert ctx.fallback_count() == f0


def test_fused_market_shapes_full_size():
    """BASELINE configs 0/1 (Market-1501 shape, D = 768 and D = 2048) at full size."""
    from daliid_b200 import synth
    for name in ("market_vit", "market_resnet50"):
        qf, gf, qp, gp, qc, gc = synth.make_config(name, device="cuda")
        a, b = _both(qf, gf, qp, gp, qc, gc, "f16x3", expect_fused=True)
        assert 0.05 < a[1] < 0.95

"""CPU tests of the oracle itself (no GPU): hand-derived known answers, agreement of the two
independently written accumulation variants and of the C restatement, an independent
cross-check against scikit-learn, and the committed golden fixtures."""
import json
import os

import numpy as np
import pytest

from oracle import c_oracle, distmat_oracle, rank_oracle as ro

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _meta():
    return json.load(open(os.path.join(GOLDEN, "meta.json")))


@pytest.mark.parametrize("kat", _meta()["kat"], ids=lambda k: k["name"])
@pytest.mark.parametrize("fn", [ro.eval_market1501_cy_f32, ro.eval_market1501_py_f64,
                                c_oracle.evaluate_rank_c], ids=["cy_f32", "py_f64", "c"])
def test_known_answers(kat, fn):
    d = np.array(kat["dist"], dtype=np.float32)
    cmc, mAP, ap, first = fn(d, kat["q_pid"], kat["g_pid"], kat["q_cam"], kat["g_cam"],
                             return_details=True)
    assert mAP == pytest.approx(kat["mAP"], abs=1e-7)
    assert list(cmc) == kat["cmc"]
    assert list(first) == kat["first_rank"]


def test_all_invalid_raises():
    d = np.zeros((2, 3), dtype=np.float32)
    for fn in (ro.eval_market1501_cy_f32, ro.eval_market1501_py_f64, c_oracle.evaluate_rank_c):
        with pytest.raises(AssertionError, match="all query identities do not appear in gallery"):
            fn(d, [1, 2], [3, 4, 5], [0, 0], [1, 1, 1])


def _random_case(seed, Q=120, G=900, ids=30, cams=5, quant=None):
    rng = np.random.default_rng(seed)
    d = rng.random((Q, G)).astype(np.float32)
    if quant:
        d = (np.round(d * quant) / quant).astype(np.float32)
    return (d, rng.integers(0, ids + 3, Q), rng.integers(0, ids, G), rng.integers(0, cams, Q),
            rng.integers(0, cams, G))


@pytest.mark.parametrize("quant", [None, 128])
def test_variants_agree(quant):
    d, qp, gp, qc, gc = _random_case(3, quant=quant)
    a = ro.eval_market1501_cy_f32(d, qp, gp, qc, gc, return_details=True)
    b = ro.eval_market1501_py_f64(d, qp, gp, qc, gc, return_details=True)
    c = c_oracle.evaluate_rank_c(d, qp, gp, qc, gc, return_details=True)
    c2 = c_oracle.evaluate_rank_c(d, qp, gp, qc, gc, tie="c_stable", return_details=True)
    # C restatement == numpy restatement of the Cython semantics, bit for bit
    assert np.array_equal(a[0], c[0]) and a[1] == c[1]
    assert np.array_equal(a[2].astype(np.float32), c[2].astype(np.float32), equal_nan=True)
    assert np.array_equal(a[3], c[3]) and np.array_equal(c[3], c2[3]) and c[1] == c2[1]
    # float32-sequential vs float64-pairwise differ only in rounding
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[3], b[3])
    assert abs(a[1] - b[1]) < 1e-6
    assert np.nanmax(np.abs(a[2] - b[2])) < 1e-6


def test_sklearn_average_precision():
    from sklearn.metrics import average_precision_score
    d, qp, gp, qc, gc = _random_case(5, Q=40, G=300, ids=10)
    _, _, ap, first = ro.eval_market1501_py_f64(d, qp, gp, qc, gc, return_details=True)
    for q in range(len(qp)):
        keep = ~((gp == qp[q]) & (gc == qc[q]))
        y = (gp[keep] == qp[q]).astype(int)
        if y.sum() == 0:
            assert first[q] == -1
            continue
        assert ap[q] == pytest.approx(average_precision_score(y, -d[q][keep].astype(np.float64)), abs=1e-12)


def test_string_labels_equal_int_labels():
    d, qp, gp, qc, gc = _random_case(7)
    a = ro.evaluate_rank(d, qp, gp, qc, gc)
    b = ro.evaluate_rank(d, qp.astype(str), gp.astype(str), qc.astype(str), gc.astype(str))
    assert np.array_equal(a[0], b[0]) and a[1] == b[1]


def test_permutation_invariance_tie_free():
    d, qp, gp, qc, gc = _random_case(11)
    perm = np.random.default_rng(0).permutation(d.shape[1])
    a = ro.evaluate_rank(d, qp, gp, qc, gc, accum="py_f64")
    b = ro.evaluate_rank(d[:, perm], qp, gp[perm], qc, gc[perm], accum="py_f64")
    assert np.array_equal(a[0], b[0]) and a[1] == b[1]
    assert np.all(np.diff(a[0]) >= 0)


@pytest.mark.parametrize("name", ["tiny", "ties", "small_gallery"])
def test_golden_rank_fixtures(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    d = z["cosine"] if name == "tiny" else z["dist"]
    for accum, fn in (("cy_f32", ro.eval_market1501_cy_f32), ("py_f64", ro.eval_market1501_py_f64)):
        cmc, mAP, ap, first = fn(d, z["q_pid"], z["g_pid"], z["q_cam"], z["g_cam"], return_details=True)
        assert np.array_equal(cmc, z[f"cmc_{accum}"])
        assert mAP == float(z[f"mAP_{accum}"])
        assert np.array_equal(ap, z[f"ap_{accum}"], equal_nan=True)
        assert np.array_equal(first, z["first_rank"])
    cmc, mAP = c_oracle.evaluate_rank_c(d, z["q_pid"], z["g_pid"], z["q_cam"], z["g_cam"])
    assert np.array_equal(cmc, z["cmc_cy_f32"]) and mAP == float(z["mAP_cy_f32"])


def test_golden_distance_fixtures():
    import torch
    z = np.load(os.path.join(GOLDEN, "tiny.npz"))
    q, g = torch.from_numpy(z["qf"]), torch.from_numpy(z["gf"])
    # CPU BLAS blocking may differ between hosts: tolerance, not bit equality
    np.testing.assert_allclose(distmat_oracle.cosine_distmat(q, g).numpy(), z["cosine"], atol=2e-6)
    np.testing.assert_allclose(distmat_oracle.sqeuclidean_distmat(q, g).numpy(), z["sqeuclidean"], rtol=1e-5)
    np.testing.assert_allclose(distmat_oracle.euclidean_distmat(q, g).numpy(), z["euclidean"], rtol=1e-5)


def test_golden_fusion_fixtures():
    import torch
    z = np.load(os.path.join(GOLDEN, "fusion.npz"))
    assert np.array_equal(distmat_oracle.fuse_mean([z["d0"], z["d1"]]), z["mean2"])
    assert np.array_equal(distmat_oracle.fuse_mean([z["d0"], z["d1"], z["d2"]]), z["mean3"])
    w = [distmat_oracle.magnitude_weights(torch.from_numpy(z[f"qm{i}"]), torch.from_numpy(z[f"gm{i}"]))
         for i in range(2)]
    assert np.array_equal(distmat_oracle.fuse_weighted(w, [z["d0"], z["d1"]]).numpy(), z["weighted"])


def test_briar_hits_match_reference_callsite():
    z = np.load(os.path.join(GOLDEN, "tiny.npz"))
    hits, top = ro.briar_rank_hits(z["cosine"], z["q_pid"], z["g_pid"])
    ref = _meta()["callsites"]["validateBRIAR.calculateMetrics"]
    assert hits == pytest.approx(ref["cmc"], abs=0)
    assert np.array_equal(top.astype(np.int32), np.load(os.path.join(GOLDEN, "tiny_top20.npy")))


def _rerank_case(Q=23, G=90, D=16, seed=0):
    import torch
    g = torch.Generator().manual_seed(seed)
    cent = torch.randn(9, D, generator=g)
    q = cent[torch.randint(0, 9, (Q,), generator=g)] + 0.7 * torch.randn(Q, D, generator=g)
    x = cent[torch.randint(0, 9, (G,), generator=g)] + 0.7 * torch.randn(G, D, generator=g)
    q = q / q.norm(dim=1, keepdim=True)
    x = x / x.norm(dim=1, keepdim=True)
    sq = lambda a, b: ((a * a).sum(1)[:, None] + (b * b).sum(1)[None, :] - 2.0 * a @ b.T).numpy()
    return (1.0 - q @ x.T).numpy(), sq(q, q), sq(x, x)


def test_rerank_oracle_invariants():
    """The restated k-reciprocal re-ranking (oracle/rerank_oracle.py) is pinned by its own
    invariants: rows of V sum to one, Jaccard self-distance of a query is 0, lambda = 1 returns the
    normalised squared input, gallery permutations permute the result."""
    from oracle import rerank_oracle as rr
    qg, qq, gg = _rerank_case()
    Q, G = qg.shape
    det = rr.re_ranking_details(qg, qq, gg, k1=8, k2=3, lambda_value=0.3)
    assert det["final"].shape == (Q, G) and det["final"].dtype == np.float32
    np.testing.assert_allclose(det["V0"].sum(1), 1.0, rtol=1e-5)
    np.testing.assert_allclose(det["V"].sum(1), 1.0, rtol=1e-5)
    assert np.all(np.abs(det["jaccard"][np.arange(Q), np.arange(Q)]) < 1e-6)
    lam1 = rr.re_ranking(qg, qq, gg, k1=8, k2=3, lambda_value=1.0)
    np.testing.assert_array_equal(lam1, det["original"][:, Q:])
    perm = np.random.default_rng(1).permutation(G)
    p = rr.re_ranking(qg[:, perm], qq, gg[np.ix_(perm, perm)], k1=8, k2=3, lambda_value=0.3)
    np.testing.assert_allclose(p, det["final"][:, perm], atol=2e-6)
    k2_1 = rr.re_ranking_details(qg, qq, gg, k1=8, k2=1, lambda_value=0.3)
    np.testing.assert_array_equal(k2_1["V"], k2_1["V0"])


def test_rerank_oracle_matches_golden():
    """Regression pin of the restated re-ranking against the committed vectors."""
    from oracle import rerank_oracle as rr
    z = np.load(os.path.join(GOLDEN, "rerank.npz"))
    for name in "abc":
        k1, k2, lam = z["params_" + name]
        out = rr.re_ranking(z["qg"], z["qq"], z["gg"], int(k1), int(k2), float(lam))
        np.testing.assert_allclose(out, z["final_" + name], rtol=0, atol=1e-7)


def test_mrfuse_oracle_matches_reference_golden():
    """oracle/mrfuse_oracle.py against the output of the reference's own libmr / Meta_Recognition
    classes (tests/golden/make_golden_mrfuse.py)."""
    from oracle import mrfuse_oracle
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "mrfuse.npz"))
    for c in "ab":
        s = [z[f"{c}_s{m}"] for m in range(3)]
        for m in range(3):
            w, fit, small = mrfuse_oracle.metarec(s[m])
            np.testing.assert_array_equal(small, z[f"{c}_small{m}"])
            np.testing.assert_allclose(fit, z[f"{c}_fit{m}"], rtol=1e-9, equal_nan=True)
            np.testing.assert_allclose(w, z[f"{c}_w{m}"], rtol=0, atol=1e-9)
        fused = mrfuse_oracle.mrfuse(*s)
        np.testing.assert_allclose(fused, z[f"{c}_fused"], rtol=0, atol=1e-9, equal_nan=True)
    # the degenerate (constant) column of case b never converges: parameters stay (0, 0), weight 0.5
    assert np.all(z["b_fit0"][5] == 0) and np.all(z["b_w0"][:, 5] == 0.5)

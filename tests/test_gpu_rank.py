"""GPU parity tests of the rank / CMC / mAP stage through the C-ABI: bit-exact against the
oracle (integer ranks, float32 and float64 accumulation) on the golden fixtures, seeded
random cases, tie stressors and the edge cases the domain has."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import rank_oracle as ro

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _check(dist, qp, gp, qc, gc, max_rank=50, where="host"):
    from daliid_b200 import metrics
    d_in = dist
    if where == "device":
        d_in = torch.from_numpy(np.ascontiguousarray(dist)).cuda()
    elif where == "device_strided":  # ld > G and a misaligned row start
        pad = torch.full((dist.shape[0], dist.shape[1] + 5), 7.0, device="cuda")
        pad[:, 1:1 + dist.shape[1]] = torch.from_numpy(np.ascontiguousarray(dist)).cuda()
        d_in = pad[:, 1:1 + dist.shape[1]]
    for accum, fn in (("cy_f32", ro.eval_market1501_cy_f32), ("py_f64", ro.eval_market1501_py_f64)):
        e_cmc, e_map, e_ap, e_first = fn(dist, qp, gp, qc, gc, max_rank, return_details=True)
        cmc, mAP, ap, first, nvalid = metrics.evaluate_rank_detailed(d_in, qp, gp, qc, gc, max_rank, accum)
        assert np.array_equal(first, e_first), accum
        assert nvalid == int((e_first > 0).sum())
        assert cmc.dtype == np.float32 and np.array_equal(cmc, e_cmc), accum
        assert np.array_equal(ap, e_ap, equal_nan=True), accum
        assert mAP == e_map, (accum, mAP, e_map)


@pytest.mark.parametrize("name", ["tiny", "ties", "small_gallery"])
@pytest.mark.parametrize("where", ["host", "device", "device_strided"])
def test_golden_fixtures(name, where):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    d = z["cosine"] if name == "tiny" else z["dist"]
    _check(d, z["q_pid"], z["g_pid"], z["q_cam"], z["g_cam"], where=where)
    # and against the committed expectations directly
    from daliid_b200 import metrics
    for accum in ("cy_f32", "py_f64"):
        cmc, mAP, ap, first, _ = metrics.evaluate_rank_detailed(
            d, z["q_pid"], z["g_pid"], z["q_cam"], z["g_cam"], 50, accum)
        assert np.array_equal(cmc, z[f"cmc_{accum}"]) and mAP == float(z[f"mAP_{accum}"])
        assert np.array_equal(ap, z[f"ap_{accum}"], equal_nan=True)
        assert np.array_equal(first, z["first_rank"])


def test_known_answers():
    from daliid_b200 import metrics
    for kat in json.load(open(os.path.join(GOLDEN, "meta.json")))["kat"]:
        d = np.array(kat["dist"], dtype=np.float32)
        cmc, mAP, ap, first, _ = metrics.evaluate_rank_detailed(
            d, kat["q_pid"], kat["g_pid"], kat["q_cam"], kat["g_cam"])
        assert list(cmc) == kat["cmc"], kat["name"]
        assert list(first) == kat["first_rank"]
        assert mAP == pytest.approx(kat["mAP"], abs=1e-7)


def _case(seed, Q, G, ids, cams, quant=None):
    rng = np.random.default_rng(seed)
    d = rng.random((Q, G)).astype(np.float32)
    if quant:
        d = (np.round(d * quant) / quant).astype(np.float32)
    return (d, rng.integers(0, ids + 2, Q), rng.integers(0, ids, G), rng.integers(0, cams, Q),
            rng.integers(0, cams, G))


@pytest.mark.parametrize("Q,G,ids,cams,quant", [
    (1, 1, 1, 2, None),            # single pair
    (3, 7, 2, 2, 4),               # rows shorter than one vector
    (33, 1023, 40, 3, None),       # G not a multiple of 4
    (64, 4097, 9, 2, 256),         # > 32 positives per query -> several threshold chunks
    (50, 3000, 1, 2, 16),          # every gallery item matches: ~1500 positives / query
    (9, 5000, 1, 1000, 8),         # > 2048 matches per query: non-staged epilogue
    (2, 70001, 300, 4, 1024),      # few queries, long rows -> column-split path
    (257, 2048, 100, 6, 65536),
])
def test_random_cases_bit_exact(Q, G, ids, cams, quant):
    _check(*_case(Q * 131 + G, Q, G, ids, cams, quant), where="device")


def test_string_labels_and_tensor_input():
    from daliid_b200 import metrics
    d, qp, gp, qc, gc = _case(5, 80, 700, 20, 4)
    e = ro.evaluate_rank(d, qp, gp, qc, gc)
    a = metrics.evaluate_rank(torch.from_numpy(d), qp.astype(str), gp.astype(str), qc.astype(str), gc.astype(str))
    assert np.array_equal(a[0], e[0]) and a[1] == e[1]
    b = metrics.evaluate_rank(d, qp, gp, qc, gc, use_cython=False)
    e2 = ro.evaluate_rank(d, qp, gp, qc, gc, use_cython=False)
    assert np.array_equal(b[0], e2[0]) and b[1] == e2[1]


def test_error_behaviour():
    from daliid_b200 import metrics
    d = np.zeros((2, 3), dtype=np.float32)
    with pytest.raises(AssertionError, match="all query identities do not appear in gallery"):
        metrics.evaluate_rank(d, [1, 2], [3, 4, 5], [0, 0], [1, 1, 1])
    with pytest.raises(NotImplementedError):
        metrics.evaluate_rank(d, [1, 2], [3, 4, 5], [0, 0], [1, 1, 1], use_metric_cuhk03=True)
    with pytest.raises(ValueError):
        metrics.evaluate_rank(d, [1], [3, 4, 5], [0, 0], [1, 1, 1])


def test_sharded_building_blocks_equal_unsharded():
    """1/2/3/8 gallery slabs emulated on one GPU through the (e) building blocks: keys and
    counts summed over slabs reproduce the single-GPU result bit for bit."""
    from daliid_b200 import metrics, sharded
    d, qp, gp, qc, gc = _case(77, 150, 5003, 37, 5, quant=512)
    e = ro.eval_market1501_cy_f32(d, qp, gp, qc, gc, return_details=True)
    dd = torch.from_numpy(d).cuda()
    ops = sharded.CudaOps()
    qpi, gpi = metrics.canonicalize_labels(qp, gp)
    qci, gci = metrics.canonicalize_labels(qc, gc)
    for world in (1, 2, 3, 8):
        plan = ops.plan(qpi, gpi, qci, gci)
        slabs = [sharded.slab_bounds(d.shape[1], world, r) for r in range(world)]
        keys = sum(ops.gather_keys(plan, dd[:, g0:g0 + gs].contiguous(), g0) for g0, gs in slabs)
        counts = sum(ops.count(plan, dd[:, g0:g0 + gs].contiguous(), g0, keys) for g0, gs in slabs)
        cmc, mAP, det = ops.finalize(plan, keys, counts, d.shape[0], d.shape[1], 50, "cy_f32")
        ops.plan_destroy(plan)
        assert np.array_equal(cmc, e[0]) and mAP == e[1], world
        assert np.array_equal(det["first_rank"], e[3])


def test_plan_cache_reuses_only_identical_labels():
    """The context keeps the last rank plan and reuses it when the labels are byte-for-byte
    equal; any change in any of the four arrays rebuilds it (results checked against the oracle
    either way)."""
    from daliid_b200 import _lib, metrics
    rng = np.random.default_rng(4)
    Q, G = 40, 500
    d = rng.random((Q, G), dtype=np.float32)
    qp = rng.integers(0, 7, Q).astype(np.int32); gp = rng.integers(0, 7, G).astype(np.int32)
    qc = rng.integers(0, 3, Q).astype(np.int32); gc = rng.integers(0, 3, G).astype(np.int32)
    ctx = _lib.get_ctx(0)
    ctx.plan_cache_enable(True)
    h0 = ctx.plan_cache_hits()
    a = metrics.evaluate_rank(d, qp, gp, qc, gc)
    b = metrics.evaluate_rank(d * 0.5 + 0.1, qp.copy(), gp.copy(), qc.copy(), gc.copy())
    assert ctx.plan_cache_hits() == h0 + 1          # second call: same labels, other distances
    e = ro.evaluate_rank(d, qp, gp, qc, gc)
    assert np.array_equal(a[0], e[0]) and a[1] == e[1]
    assert np.array_equal(b[0], e[0]) and b[1] == e[1]   # monotone map of d: same ranking
    for arr in (qp, gp, qc, gc):                    # one changed label anywhere: rebuilt
        old = arr[-1]
        arr[-1] = old + 1
        h = ctx.plan_cache_hits()
        r = metrics.evaluate_rank(d, qp, gp, qc, gc)
        assert ctx.plan_cache_hits() == h
        e = ro.evaluate_rank(d, qp, gp, qc, gc)
        assert np.array_equal(r[0], e[0]) and r[1] == e[1]
        arr[-1] = old
    ctx.plan_cache_enable(False)
    h = ctx.plan_cache_hits()
    metrics.evaluate_rank(d, qp, gp, qc, gc); metrics.evaluate_rank(d, qp, gp, qc, gc)
    assert ctx.plan_cache_hits() == h
    ctx.plan_cache_enable(True)


@pytest.mark.parametrize("Q,G,k1,k2,lam", [(23, 90, 8, 3, 0.3), (64, 700, 20, 6, 0.3), (40, 333, 28, 1, 0.5),
                                           (5, 40, 4, 2, 0.0)])
def test_re_ranking_matches_oracle(Q, G, k1, k2, lam):
    """k-reciprocal re-ranking (torchreid.utils.re_ranking's signature) against the
    restated numpy form.  Neighbour sets are integer work (a wrong member would move a value by
    ~1e-2); values agree to fp32 rounding -- exp() differs in the last ulp between numpy and CUDA."""
    import torch
    from daliid_b200 import metrics
    from oracle import rerank_oracle as rr
    g = torch.Generator().manual_seed(Q + G)
    D = 24
    cent = torch.randn(11, D, generator=g)
    q = cent[torch.randint(0, 11, (Q,), generator=g)] + 0.6 * torch.randn(Q, D, generator=g)
    x = cent[torch.randint(0, 11, (G,), generator=g)] + 0.6 * torch.randn(G, D, generator=g)
    x[7] = x[3]  # exact duplicates: ties in the neighbour lists
    q = q / q.norm(dim=1, keepdim=True)
    x = x / x.norm(dim=1, keepdim=True)
    sq = lambda a, b: ((a * a).sum(1)[:, None] + (b * b).sum(1)[None, :] - 2.0 * a @ b.T)
    qg, qq, gg = (1.0 - q @ x.T), sq(q, q), sq(x, x)
    exp = rr.re_ranking(qg.numpy(), qq.numpy(), gg.numpy(), k1, k2, lam)
    out = metrics.re_ranking(qg.numpy(), qq.numpy(), gg.numpy(), k1, k2, lam)
    assert out.dtype == np.float32 and out.shape == (Q, G)
    assert np.abs(out - exp).max() <= 2e-6
    out_d = metrics.re_ranking(qg.cuda(), qq.cuda(), gg.cuda(), k1, k2, lam)
    assert out_d.is_cuda and np.array_equal(out_d.cpu().numpy(), out)
    if lam == 0.0:
        return
    # the re-ranked matrix feeds the evaluator like any other
    qp = np.arange(Q, dtype=np.int32) % 11
    gp = np.arange(G, dtype=np.int32) % 11
    a = metrics.evaluate_rank(out, qp, gp, np.zeros(Q, np.int32), np.ones(G, np.int32))
    b = ro.evaluate_rank(out, qp, gp, np.zeros(Q, np.int32), np.ones(G, np.int32))
    assert np.array_equal(a[0], b[0]) and a[1] == b[1]


def test_re_ranking_golden():
    from daliid_b200 import metrics
    z = np.load(os.path.join(GOLDEN, "rerank.npz"))
    for name in "abc":
        k1, k2, lam = z["params_" + name]
        out = metrics.re_ranking(z["qg"], z["qq"], z["gg"], int(k1), int(k2), float(lam))
        assert np.abs(out - z["final_" + name]).max() <= 2e-6, name


def test_property_random_cases_hypothesis():
    """Property test (hypothesis, derandomised): arbitrary small shapes, label alphabets, value
    alphabets full of ties / signed zeros / infinities / NaN, any max_rank and both accumulation
    modes -- CMC, AP, first ranks and mAP equal the oracle bit for bit; queries with no valid match
    are skipped; all-invalid inputs raise torchreid's AssertionError."""
    from hypothesis import given, settings, strategies as st, HealthCheck
    from daliid_b200 import metrics

    specials = [0.0, -0.0, 0.25, 0.5, 0.5, 1.0, 1.5, np.inf, -np.inf, np.nan, 1e-38, -3.0, 7.0]

    @st.composite
    def cases(draw):
        Q = draw(st.integers(1, 12))
        G = draw(st.integers(1, 90))
        n_ids = draw(st.integers(1, 6))
        n_cams = draw(st.integers(1, 3))
        seed = draw(st.integers(0, 2 ** 31 - 1))
        mode = draw(st.sampled_from(["specials", "quantised", "continuous"]))
        max_rank = draw(st.sampled_from([1, 3, 20, 50]))
        rng = np.random.default_rng(seed)
        if mode == "specials":
            d = rng.choice(np.array(specials, dtype=np.float32), size=(Q, G))
        elif mode == "quantised":
            d = (rng.integers(0, 5, size=(Q, G)) / 4).astype(np.float32)
        else:
            d = rng.random((Q, G), dtype=np.float32) * 2 - 0.5
        return (d.astype(np.float32), rng.integers(0, n_ids, Q).astype(np.int32),
                rng.integers(0, n_ids, G).astype(np.int32), rng.integers(0, n_cams, Q).astype(np.int32),
                rng.integers(0, n_cams, G).astype(np.int32), max_rank)

    @settings(max_examples=120, deadline=None, derandomize=True,
              suppress_health_check=[HealthCheck.too_slow, HealthCheck.data_too_large])
    @given(cases())
    def run(case):
        d, qp, gp, qc, gc, max_rank = case
        for accum, fn in (("cy_f32", ro.eval_market1501_cy_f32), ("py_f64", ro.eval_market1501_py_f64)):
            try:
                e_cmc, e_map, e_ap, e_first = fn(d, qp, gp, qc, gc, max_rank, return_details=True)
            except AssertionError:
                with pytest.raises(AssertionError):
                    metrics.evaluate_rank_detailed(d, qp, gp, qc, gc, max_rank, accum)
                continue
            cmc, mAP, ap, first, _ = metrics.evaluate_rank_detailed(d, qp, gp, qc, gc, max_rank, accum)
            assert np.array_equal(first, e_first)
            assert np.array_equal(cmc, e_cmc) and mAP == e_map
            assert np.array_equal(ap, e_ap, equal_nan=True)

    run()


def test_byte_counter_rows_at_the_boundary():
    """the single-launch byte-counter path (more than 64 positives per query) must not
    wrap its 8-bit private counters.  G in (64512, 65280] gives threads 0..191 256 elements; with a
    worst-ranked positive nearly the whole row falls into one bucket.  Such rows now take the split
    launch; G = 64000 is the widest single-launch row.  Both are checked against the oracle."""
    from daliid_b200 import metrics
    rng = np.random.default_rng(7)
    for G in (64000, 65280):
        Q, npos = 3, 90
        d = rng.random((Q, G), dtype=np.float32) * 0.5          # negatives: 0 .. 0.5
        g_pid = np.full(G, 1000, dtype=np.int32)
        g_cam = np.ones(G, dtype=np.int32)
        q_pid = np.arange(Q, dtype=np.int32)
        q_cam = np.zeros(Q, dtype=np.int32)
        for q in range(Q):
            cols = rng.choice(G, npos, replace=False)
            g_pid[cols] = q                                     # positives of query q ...
            d[q, cols] = 0.9 + 0.001 * rng.random(npos, dtype=np.float32)   # ... ranked last
        e = ro.eval_market1501_cy_f32(d, q_pid, g_pid, q_cam, g_cam, 50, return_details=True)
        cmc, mAP, ap, first, _ = metrics.evaluate_rank_detailed(torch.from_numpy(d).cuda(), q_pid, g_pid,
                                                                 q_cam, g_cam)
        assert np.array_equal(first, e[3]) and mAP == e[1] and np.array_equal(cmc, e[0]), G


def test_two_streams_share_the_context_safely():
    """the context's workspaces are ordered on the stream of the previous call.  A second
    call issued under another torch stream must wait for it (dali_ctx_set_stream hands over with an
    event): the first matrix stays intact although the second call reuses the operand planes."""
    from daliid_b200 import metrics
    g = torch.Generator().manual_seed(11)
    q = torch.randn(3000, 1024, generator=g).cuda()
    g1 = torch.randn(9000, 1024, generator=g).cuda()
    g2 = torch.randn(9000, 1024, generator=g).cuda()
    ref1 = metrics.compute_distance_matrix(q, g1, "cosine").clone()
    ref2 = metrics.compute_distance_matrix(q, g2, "cosine").clone()
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for _ in range(5):
        with torch.cuda.stream(s1):
            d1 = metrics.compute_distance_matrix(q, g1, "cosine")
        with torch.cuda.stream(s2):
            d2 = metrics.compute_distance_matrix(q, g2, "cosine")
        torch.cuda.synchronize()
        assert torch.equal(d1, ref1) and torch.equal(d2, ref2)


def test_entry_points_restore_the_callers_device():
    """an entry point switches to its context's device and puts the caller's back."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from daliid_b200 import metrics
    torch.cuda.set_device(0)
    x = torch.randn(64, 128, device="cuda:1")
    metrics.compute_distance_matrix(x, x, "cosine")
    assert torch.cuda.current_device() == 0
    assert torch.empty(1, device="cuda").device.index == 0


# ---- counting kernel v3 (FMA bins + bit-sliced counters; 8-bit counters for 65..254 positives) --------

def _labels(Q, G, per_id, ncam, rng):
    """per_id gallery items per identity, cameras random: positives per query up to per_id."""
    n_ids = max(1, G // per_id)
    gp = np.arange(G) % n_ids
    gc = rng.integers(0, ncam, G)
    qp = rng.integers(0, n_ids, Q)
    qc = rng.integers(0, ncam, Q)
    return qp.astype(np.int32), gp.astype(np.int32), qc.astype(np.int32), gc.astype(np.int32)


@pytest.mark.parametrize("per_id", [1, 2, 31, 32, 33, 63, 64, 65, 66, 130, 254])
@pytest.mark.parametrize("kind", ["cosine_like", "clustered", "wide_range", "negative", "constant", "quantised"])
def test_v3_threshold_count_boundaries_and_value_distributions(per_id, kind):
    """1 / 32 / 33 / 64 / 65 / 254 thresholds per query (one and two mask words, the 8-bit counters) under
    value distributions that stress the bin map: thresholds much closer together than their size
    (the scale cap), twelve decades of dynamic range, negative values, one repeated value (every
    element shares the thresholds' bin), heavily quantised values (ties decided by gallery id)."""
    rng = np.random.default_rng(per_id * 7 + len(kind))
    Q, G = 9, 1531 if per_id < 100 else 2047
    qp, gp, qc, gc = _labels(Q, G, per_id, 4, rng)
    if kind == "cosine_like":
        d = rng.random((Q, G), dtype=np.float32) * 1.4 + 0.05
    elif kind == "clustered":
        d = (1.0 + rng.random((Q, G)) * 3e-6).astype(np.float32)
    elif kind == "wide_range":
        d = (10.0 ** rng.uniform(-18, 18, (Q, G))).astype(np.float32)
    elif kind == "negative":
        d = (rng.standard_normal((Q, G)) * 50 - 100).astype(np.float32)
    elif kind == "constant":
        d = np.full((Q, G), 0.625, dtype=np.float32)
        d[:, ::7] = 0.5
    else:
        d = (np.round(rng.random((Q, G)) * 8) / 8).astype(np.float32)
    for where in ("device", "device_strided"):
        _check(d, qp, gp, qc, gc, where=where)


def test_v3_non_finite_rows_take_the_generic_path():
    rng = np.random.default_rng(5)
    Q, G = 8, 700
    qp, gp, qc, gc = _labels(Q, G, 20, 3, rng)
    d = rng.random((Q, G), dtype=np.float32)
    d[0, :] = np.nan                      # every threshold NaN
    d[1, gp == qp[1]] = np.inf            # thresholds +inf, other elements finite
    d[2, gp == qp[2]] = -np.inf
    d[3, ::3] = np.nan                    # NaN elements among finite thresholds (and some NaN thresholds)
    d[4, 5] = np.inf; d[4, 6] = -np.inf; d[4, 7] = -0.0; d[4, 8] = 0.0
    d[5, :] = 0.0; d[5, ::2] = -0.0       # +-0 everywhere
    d[6, :] = np.float32(1e-45)           # denormals
    d[6, ::5] = np.float32(3e-45)
    _check(d, qp, gp, qc, gc, where="device")


def test_v3_row_lengths_around_the_vector_loop():
    """Head / tail columns of misaligned rows and rows shorter than one iteration of the ring."""
    rng = np.random.default_rng(9)
    for G in (1, 2, 3, 4, 5, 63, 64, 65, 511, 512, 513, 2047, 2048, 2049, 4100):
        Q = 5
        qp, gp, qc, gc = _labels(Q, G, min(G, 6), 2, rng)
        d = rng.random((Q, G), dtype=np.float32)
        if not (gp[None, :] == qp[:, None]).any():
            continue
        try:
            _check(d, qp, gp, qc, gc, where="device_strided")
        except AssertionError as e:
            if "do not appear in gallery" in str(e):  # every match same camera: upstream raises too
                continue
            raise

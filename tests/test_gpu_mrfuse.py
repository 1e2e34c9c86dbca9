"""Meta-recognition fusion on the GPU against the reference's own output
(tests/golden/mrfuse.npz, produced by running the reference's libmr / Meta_Recognition classes)
and against the CPU restatement in oracle/mrfuse_oracle.py."""
import os

import numpy as np
import pytest
import torch

from daliid_b200 import evaluate, metrics
from oracle import mrfuse_oracle

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "mrfuse.npz")
# the reference's fp32 log / mean intermediates move the fitted shape by ~1e-7 relative between
# libm implementations; the fused similarity is compared at the path's 1e-5 tolerance
FIT_RTOL, W_ATOL, FUSED_ATOL = 2e-6, 1e-5, 1e-5


def _check_case(z, c, device):
    s = [z[f"{c}_s{m}"] for m in range(3)]
    if device:
        s = [torch.from_numpy(x).cuda() for x in s]
    fused, det = metrics.mrfuse(s, 20, False, 1.0, return_details=True)
    fused = fused.cpu().numpy() if device else fused
    w = det["weights"].cpu().numpy() if device else det["weights"]
    assert fused.dtype == np.float64 and w.dtype == np.float64
    for m in range(3):
        np.testing.assert_array_equal(det["small"][m], z[f"{c}_small{m}"])
        fit, ref = det["fit"][m], z[f"{c}_fit{m}"]
        np.testing.assert_array_equal(np.isnan(fit), np.isnan(ref))
        np.testing.assert_array_equal(fit == 0, ref == 0)
        ok = ~np.isnan(ref)
        np.testing.assert_allclose(fit[ok], ref[ok], rtol=FIT_RTOL, atol=0)
        np.testing.assert_allclose(w[m], z[f"{c}_w{m}"], rtol=0, atol=W_ATOL)
    ref = z[f"{c}_fused"]
    np.testing.assert_array_equal(np.isnan(fused), np.isnan(ref))
    np.testing.assert_allclose(fused[~np.isnan(ref)], ref[~np.isnan(ref)], rtol=0, atol=FUSED_ATOL)


@pytest.mark.parametrize("case", ["a", "b"])
@pytest.mark.parametrize("device", [False, True])
def test_mrfuse_matches_reference_golden(case, device):
    _check_case(np.load(GOLD), case, device)


def _scores(seed, Q, G, D, ids=64, sigma=2.0):
    g = torch.Generator().manual_seed(seed)
    c = torch.randn(ids, D, generator=g)
    q = c[torch.randint(0, ids, (Q,), generator=g)] + sigma * torch.randn(Q, D, generator=g)
    ga = c[torch.randint(0, ids, (G,), generator=g)] + sigma * torch.randn(G, D, generator=g)
    q = q / q.norm(dim=1, keepdim=True)
    ga = ga / ga.norm(dim=1, keepdim=True)
    return q @ ga.T


@pytest.mark.parametrize("use_columns,killscale", [(False, 1.0), (True, 1.0), (False, 0.5)])
def test_metarec_vs_oracle_medium(use_columns, killscale):
    Q, G = 700, 333
    s = _scores(5, Q, G, 64)
    if not use_columns and killscale == 1.0:
        w_ref, fit_ref, small_ref = mrfuse_oracle.metarec(s)
    else:
        # the oracle spells out the mrfuse configuration only; the two other metarec variants differ
        # in which entries are reduced before the fit -- restated here with the oracle's pieces
        t = s.clone()
        src = t.T.contiguous() if use_columns else t
        tv, ti = torch.topk(src, 20, dim=1)
        src = src - killscale * torch.zeros_like(src).scatter_(1, ti, tv)
        cols = torch.nan_to_num(src if use_columns else src.T, 0)
        tail = Q - 21
        srt = torch.topk(cols, tail, dim=1).values
        small = srt[:, tail - 1]
        fit_ref = mrfuse_oracle.weibull_fit(srt + 1 - small[:, None])
        d = (s + 1 - small[None, :]).clamp(min=0)
        w_ref = torch.nan_to_num(mrfuse_oracle.weibull_cdf(d, fit_ref[:, 1][None, :], fit_ref[:, 0][None, :]), 0).numpy()
        small_ref = small.numpy()
    _, det = metrics.mrfuse([s.cuda()], 20, use_columns, killscale, return_details=True)
    np.testing.assert_array_equal(det["small"][0], small_ref)
    np.testing.assert_allclose(det["fit"][0], fit_ref, rtol=FIT_RTOL)
    np.testing.assert_allclose(det["weights"][0].cpu().numpy(), w_ref, rtol=0, atol=W_ATOL)


def test_mrfuse_class_and_properties():
    Q, G = 300, 200
    s = [_scores(20 + m, Q, G, 48) for m in range(3)]
    fused = evaluate.Meta_Recognition().mrfuse(*[x.cuda() for x in s])
    ref = mrfuse_oracle.mrfuse(*s)
    assert isinstance(fused, np.ndarray) and fused.dtype == np.float64 and fused.shape == (Q, G)
    np.testing.assert_allclose(fused, ref, rtol=0, atol=FUSED_ATOL)
    # a weighted mean with non-negative weights stays inside the per-element range of its inputs
    st = np.stack([x.numpy().astype(np.float64) for x in s])
    assert np.all(fused <= st.max(0) + 1e-12) and np.all(fused >= st.min(0) - 1e-12)
    # fusing a matrix with itself returns it
    same = metrics.mrfuse([s[0].cuda(), s[0].cuda()], 20).cpu().numpy()
    np.testing.assert_allclose(same, s[0].numpy().astype(np.float64), rtol=1e-15, atol=0)
    # host inputs give the same bits as device inputs
    host = metrics.mrfuse([x.numpy() for x in s], 20)
    np.testing.assert_array_equal(host, fused)


def test_mrfuse_recompute_path_equals_staged(monkeypatch):
    """Columns too long for shared-memory staging recompute ln(x) every Newton step: same fit."""
    s = [_scores(60 + m, 500, 90, 32).cuda() for m in range(2)]
    a, da = metrics.mrfuse(s, 20, return_details=True)
    monkeypatch.setenv("DALI_MRFUSE_STAGED", "0")
    b, db = metrics.mrfuse(s, 20, return_details=True)
    np.testing.assert_array_equal(da["small"], db["small"])
    np.testing.assert_allclose(da["fit"], db["fit"], rtol=1e-12)
    np.testing.assert_allclose(a.cpu().numpy(), b.cpu().numpy(), rtol=1e-12, atol=0)


def test_mrfuse_argument_errors():
    s = _scores(1, 21, 40, 8).cuda()
    with pytest.raises(Exception):
        metrics.mrfuse([s], 20)  # tail would be empty
    with pytest.raises(ValueError):
        metrics.mrfuse([], 20)


def test_mrfuse_full_size_properties():
    """C1 shape: three 3368 x 15913 matrices; checked through properties + a column sample against the oracle."""
    Q, G, D = 3368, 15913, 96
    s = [_scores(40 + m, Q, G, D, ids=751, sigma=2.5).cuda() for m in range(3)]
    fused, det = metrics.mrfuse(s, 20, return_details=True)
    st = torch.stack([x.double() for x in s])
    assert bool(((fused <= st.max(0).values + 1e-12) & (fused >= st.min(0).values - 1e-12)).all())
    cols = np.arange(0, G, 997)
    for m in range(3):
        sub = s[m].cpu()
        srt, small = mrfuse_oracle.tail_of_columns(sub)
        fit_ref = mrfuse_oracle.weibull_fit(srt[cols] + 1 - small[cols, None])
        np.testing.assert_array_equal(det["small"][m][cols], small[cols].numpy())
        np.testing.assert_allclose(det["fit"][m][cols], fit_ref, rtol=FIT_RTOL)

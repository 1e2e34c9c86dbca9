"""GPU tests of the drop-in entry points: same signatures, return values and printed text as
the reference's call sites (the expected stdout in tests/golden/meta.json was captured by
running the reference's own functions, lifted unmodified from /root/reference)."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _rows(pid, cam):
    return np.array([["img_%05d.jpg" % i, str(int(pid[i])), str(int(cam[i])), "person"]
                     for i in range(len(pid))])


@pytest.fixture(scope="module")
def tiny():
    z = np.load(os.path.join(GOLDEN, "tiny.npz"))
    meta = json.load(open(os.path.join(GOLDEN, "meta.json")))["callsites"]
    return z, _rows(z["q_pid"], z["q_cam"]), _rows(z["g_pid"], z["g_cam"]), meta


def test_validateModels_calculateMetrics(tiny, capsys):
    from daliid_b200.validateModels import validationManager
    z, queries, gallery, meta = tiny
    v = validationManager.getValidator("Market")
    cmc, mAP = v.calculateMetrics(torch.from_numpy(z["cosine"]), queries, gallery)
    ref = meta["validateModels.calculateMetrics[cy_f32]"]
    assert capsys.readouterr().out == ref["stdout"]
    assert mAP == ref["mAP"] and [float(x) for x in cmc] == ref["cmc"]


def test_calculate_metrics_scripts(tiny, capsys):
    from daliid_b200 import evaluate as ev
    z, queries, gallery, meta = tiny
    assert ev.calculate_metrics(z["cosine"], queries, gallery) is None
    assert capsys.readouterr().out == meta["evaluate.calculate_metrics[cy_f32]"]["stdout"]
    ev.calculateMetrics(queries, gallery, torch.from_numpy(z["cosine"]))
    assert capsys.readouterr().out == meta["evaluateCleanATModels.calculateMetrics[cy_f32]"]["stdout"]


def test_validateBRIAR(tiny, capsys):
    from daliid_b200.validateModels import validationManager
    z, queries, gallery, meta = tiny
    v = validationManager.getValidator("BRIAR")
    cmc, zero = v.calculateMetrics(torch.from_numpy(z["cosine"]), queries, gallery)
    ref = meta["validateBRIAR.calculateMetrics"]
    assert capsys.readouterr().out == ref["stdout"]
    assert [float(x) for x in cmc] == ref["cmc"] and zero == 0


class _FakeModel:
    def eval(self):
        return self


def test_validate_end_to_end(tiny, capsys):
    """validateModels.validate(queries, gallery, model) with a stub extractor: same return
    triple (cmc, mAP, distmat) and the same report as the reference on the same features."""
    from daliid_b200.validateModels import validateModels
    z, queries, gallery, meta = tiny
    feats = {len(queries): torch.from_numpy(z["qf"]), len(gallery): torch.from_numpy(z["gf"])}
    v = validateModels()
    v.feature_extractor = staticmethod(lambda subset, h, w, model, bs, gpu: feats[len(subset)])
    v.setParameters(256, 128, False, 0)
    cmc, mAP, distmat = v.validate(queries, gallery, _FakeModel())
    out = capsys.readouterr().out
    assert out == meta["validateModels.calculateMetrics[cy_f32]"]["stdout"]
    assert isinstance(distmat, torch.Tensor) and tuple(distmat.shape) == (len(queries), len(gallery))
    np.testing.assert_allclose(distmat.cpu().numpy(), z["cosine"], atol=1e-5)
    assert abs(mAP - meta["validateModels.calculateMetrics[cy_f32]"]["mAP"]) < 1e-4


def test_ensemble_scripts(tiny, capsys):
    from daliid_b200 import evaluate as ev
    from oracle import distmat_oracle as do
    z, queries, gallery, _ = tiny
    q1, g1 = torch.from_numpy(z["qf"]), torch.from_numpy(z["gf"])
    q2, g2 = q1.flip(1).contiguous() * 1.5, g1.flip(1).contiguous() * 0.5
    d1, d2, de = ev.evaluate_ensembled_models(q1, g1, q2, g2, queries, gallery)
    assert np.array_equal(de, do.fuse_mean([d1, d2]))
    mags = [torch.norm(t, dim=1, keepdim=True) for t in (q1, g1, q2, g2)]
    out = ev.evaluate_clean_at_models(q1, g1, q2, g2, queries, gallery, magnitudes=mags)
    w = [do.magnitude_weights(mags[0], mags[1]), do.magnitude_weights(mags[2], mags[3])]
    assert np.array_equal(out["weighted"], do.fuse_weighted(w, [out["clean"], out["distortion"]]).numpy())
    assert capsys.readouterr().out.count("Computing CMC and mAP ...") == 3 + 5


def test_msmt17_validator_balanced_accuracy(capsys):
    from daliid_b200.validateModels import MSMT17_validator
    rng = np.random.default_rng(4)
    ids = rng.integers(0, 9, 200)
    centers = rng.standard_normal((9, 48)).astype(np.float32)
    tr = torch.from_numpy(centers[ids] + 0.4 * rng.standard_normal((200, 48)).astype(np.float32))
    vids = rng.integers(0, 9, 90)
    va = torch.from_numpy(centers[vids] + 0.4 * rng.standard_normal((90, 48)).astype(np.float32))
    rows = lambda i: np.array([["x", str(int(v)), "0", "person"] for v in i])
    trainer = type("T", (), dict(img_height=1, img_width=1, gpu_indexes=[0], model_name="m", version="v"))()
    v = MSMT17_validator(rows(ids), rows(vids), trainer, ".")
    acc = v.balanced_accuracy_from_features(tr, va)
    # reference arithmetic (validateModels.py:159-195) on CPU
    s = tr / torch.norm(tr, dim=1, keepdim=True)
    labels = np.unique(ids)
    cs = torch.cat([torch.mean(s[ids == l], dim=0, keepdim=True) for l in labels])
    cs = cs / torch.norm(cs, dim=1, keepdim=True)
    S = torch.mm(va / torch.norm(va, dim=1, keepdim=True), cs.T)
    top = labels[torch.topk(S, k=5, dim=1, largest=True).indices.numpy()][:, 0]
    tm = vids == top
    ref = np.mean([tm[vids == l].mean() for l in np.unique(vids)])
    assert acc == pytest.approx(ref, abs=1e-12)


def test_device_feature_sink_equals_reference_loop():
    """the device-resident sink returns exactly what the reference's host loop
    (getFeatures.py:56-67: model(batch.cuda()).cpu() + torch.cat) returns, and feeds
    validate()'s fused call without a host round trip."""
    from torch.utils.data import DataLoader, TensorDataset
    from daliid_b200 import getFeatures as gf_mod, metrics
    g = torch.Generator().manual_seed(7)
    x = torch.randn(1037, 24, generator=g)
    model = torch.nn.Sequential(torch.nn.Linear(24, 96), torch.nn.ReLU(), torch.nn.Linear(96, 40)).cuda()
    loader = DataLoader(TensorDataset(x), batch_size=100)
    dev = gf_mod.extract_features_to_device(loader, model, 0)
    ref = []
    model.eval()
    with torch.no_grad():
        for (batch,) in loader:
            fvs = model(batch.cuda(0)).data.cpu()
            ref = fvs if len(ref) == 0 else torch.cat((ref, fvs), 0)
    assert dev.is_cuda and dev.shape == (1037, 40) and torch.equal(dev.cpu(), ref)
    qp = np.arange(37, dtype=np.int32) % 9
    gp = np.arange(1000, dtype=np.int32) % 9
    qc = np.zeros(37, dtype=np.int32)
    gc = np.ones(1000, dtype=np.int32)
    a = metrics.evaluate_features(dev[:37].contiguous(), dev[37:].contiguous(), qp, gp, qc, gc)
    b = metrics.evaluate_features(ref[:37].contiguous(), ref[37:].contiguous(), qp, gp, qc, gc)
    assert np.array_equal(a[0], b[0]) and a[1] == b[1]


def test_training_side_proxy_utilities():
    """the trainer's two torch.cdist uses (train_encodersKIT.py:146-153, 252-284) on
    the library's Euclidean kernel -- same numbers as the reference expressions on the CPU."""
    from daliid_b200 import proxies
    g = torch.Generator().manual_seed(2)
    X = torch.randn(400, 96, generator=g)
    X = X / torch.norm(X, dim=1, keepdim=True)
    labels = np.arange(400) // 5
    # reference: min distance between different-label proxies
    d = torch.cdist(X, X, p=2.0)
    temp = np.array([labels])
    rep = temp.repeat(temp.shape[1], axis=0)
    mask = torch.Tensor(np.int32(rep == rep.T))
    ref_min = torch.min(mask * torch.max(d) + (1 - mask) * d).item()
    assert abs(proxies.min_negative_distance(X, labels) - ref_min) <= 1e-5
    # reference: farthest-point selection with the same random start
    np.random.seed(12)
    sel, mx = proxies.selectProxiesByTriagulation(X[:60], num_proxies=5)
    np.random.seed(12)
    cum = torch.ones(60) * torch.max(d[:60, :60])
    ref = [np.random.choice(60)]
    for j in range(4):
        cum = np.minimum(cum, d[:60, :60][ref[j]])
        ref.append(torch.argsort(cum, stable=True)[-1].item())
    assert sel.tolist() == ref
    assert abs(mx - torch.max(d[:60, :60][ref, :][:, ref]).item()) <= 1e-5


def test_roc_over_all_pairs_equals_sklearn(tiny, tmp_path, monkeypatch):
    """the reference's ROC branch (evaluateCleanATModels.py:276-292) -- same arrays
    as sklearn.metrics.roc_curve on the flattened pair labels / scores, ties included."""
    from sklearn.metrics import roc_curve
    from daliid_b200 import evaluate, verification
    z, q_rows, g_rows, _ = tiny
    rng = np.random.default_rng(0)
    for dist in (z["cosine"], (np.round(rng.random((60, 500)) * 64) / 32).astype(np.float32)):
        Q, G = dist.shape
        qp = z["q_pid"][:Q] if Q <= len(z["q_pid"]) else rng.integers(0, 9, Q)
        gp = z["g_pid"][:G] if G <= len(z["g_pid"]) else rng.integers(0, 9, G)
        labels = np.int32(np.asarray(qp)[:, None] == np.asarray(gp)[None, :]).flatten()
        preds = 1.0 - dist.flatten() / 2.0
        e_fpr, e_tpr, e_thr = roc_curve(labels, preds, pos_label=1)
        fpr, tpr, thr = verification.roc_curve_pairs(dist, qp, gp)
        assert np.array_equal(fpr, e_fpr) and np.array_equal(tpr, e_tpr) and np.array_equal(thr, e_thr)
    monkeypatch.chdir(tmp_path)
    evaluate.calculateMetrics(q_rows, g_rows, z["cosine"], pooling="gap", version="t")
    assert np.array_equal(np.load(tmp_path / "FPR_t.npy"), roc_curve(
        np.int32(z["q_pid"][:, None] == z["g_pid"][None, :]).flatten(), 1.0 - z["cosine"].flatten() / 2.0)[0])


# ---- whole ranked lists (get_subset*) -------------------------------------------

@pytest.mark.parametrize("Q,G", [(1, 100003), (7, 4099), (300, 257), (1, 1), (3, 4096), (2, 4097), (70000, 5),
                                 (129, 12936)])
@pytest.mark.parametrize("descending", [False, True])
def test_argsort_rows_equals_torch_stable(Q, G, descending):
    from daliid_b200 import metrics
    g = torch.Generator().manual_seed(Q * 31 + G)
    d = torch.randn(Q, G, generator=g)
    d = torch.round(d * 16) / 16  # heavy ties: the tie order must be ascending column
    if G > 10:
        d[0, 3] = float("nan"); d[0, 5] = float("inf"); d[0, 7] = -float("inf"); d[0, 8] = -0.0; d[0, 9] = 0.0
    ref = torch.argsort(d, dim=1, descending=descending, stable=True)
    out = metrics.argsort_rows(d.cuda(), descending)
    assert out.dtype == torch.int32 and out.is_cuda
    assert torch.equal(out.cpu().long(), ref)
    host = metrics.argsort_rows(d.numpy(), descending)
    np.testing.assert_array_equal(host, ref.numpy())


def test_argsort_rows_strided_input_and_distances_in_place():
    """A row-padded matrix (what compute_distance_matrix returns) and a slice of it."""
    from daliid_b200 import metrics
    g = torch.Generator().manual_seed(11)
    base = torch.randn(64, 1000, generator=g).cuda()
    view = base[:, 3:903]  # leading dimension 1000, 900 columns, unaligned start
    out = metrics.argsort_rows(view, False)
    assert torch.equal(out.cpu().long(), torch.argsort(view.cpu(), dim=1, stable=True))


def test_rank_by_similarity_matches_reference_arithmetic():
    """get_subset / get_subset_one_encoder (getFeatures.py:243-353) from features."""
    from daliid_b200.getFeatures import rank_by_similarity
    g = torch.Generator().manual_seed(3)
    N, D = 20011, 256
    centers = torch.randn(40, D, generator=g)
    lab = torch.randint(0, 40, (N,), generator=g)
    trains = [centers[lab] + 1.5 * torch.randn(N, D, generator=g) for _ in range(3)]
    sels = [centers[7:8] + 1.5 * torch.randn(1, D, generator=g) for _ in range(3)]

    def ref_sim(sel, tr):
        sel = sel / torch.norm(sel, dim=1, keepdim=True)
        tr = tr / torch.norm(tr, dim=1, keepdim=True)
        return torch.mm(sel, tr.T)

    sim1 = ref_sim(sels[0], trains[0])
    sim3 = (sim1 + ref_sim(sels[1], trains[1]) + ref_sim(sels[2], trains[2])) / 3
    for sim, order in ((sim1, rank_by_similarity(sels[0].cuda(), trains[0].cuda())),
                       (sim3, rank_by_similarity([s.cuda() for s in sels], [t.cuda() for t in trains]))):
        order = order.cpu()
        assert order.dtype == torch.int64 and sorted(order.tolist()) == list(range(N))
        # the reference's own ordering, up to swaps of samples whose similarities differ by rounding only
        ref = torch.argsort(sim, dim=1, descending=True)[0]
        got = sim[0][order]
        assert bool((got[:-1] - got[1:] >= -2e-6).all())
        topK = int(N * 0.1)
        assert len(set(order[:topK].tolist()) ^ set(ref[:topK].tolist())) <= 4
        assert (lab[order[:200]] == 7).float().mean() > 0.9


def test_roc_histogram_kernel_lies_on_the_exact_curve():
    """csrc/roc.cu: the binned ROC (one streaming pass, no sort) against (1) numpy's histogram of the
    same fp32 scores -- bit-exact counts -- and (2) the exact curve of the stock formulation that
    equals scikit-learn's: every binned point is a point of the exact step curve."""
    from daliid_b200 import metrics, synth, verification
    qf, gf, qp, gp, qc, gc = synth.make_config("small", device="cuda")
    d = metrics.compute_distance_matrix(qf, gf, "cosine")
    bins = 4096
    fpr, tpr, thr = verification.roc_curve_binned(d, qp, gp, bins=bins)
    dn = d.cpu().numpy()
    score = (np.float32(1.0) - dn * np.float32(0.5)).astype(np.float32)
    y = (qp[:, None] == gp[None, :])
    b = np.clip(np.floor(score.astype(np.float32) * np.float32(bins)).astype(np.int64), 0, bins - 1)
    pos = np.bincount(b[y], minlength=bins)
    neg = np.bincount(b[~y], minlength=bins)
    assert np.array_equal(np.cumsum(pos[::-1]) / max(pos.sum(), 1), tpr[1:])
    assert np.array_equal(np.cumsum(neg[::-1]) / max(neg.sum(), 1), fpr[1:])
    assert fpr[0] == 0 and tpr[0] == 0 and fpr[-1] == 1 and tpr[-1] == 1 and np.isinf(thr[0])
    # on the exact curve: at the smallest score of a bin, the exact (fpr, tpr) equals the binned point
    efpr, etpr, ethr = verification.roc_curve_pairs(d, qp, gp)
    for i in (bins // 2 + 7, bins // 2 + 40, bins // 2 + 200):
        sel = b >= bins - i
        if not sel.any():
            continue
        t = score[sel].min()
        exact_tpr = (y & (score >= t)).sum() / y.sum()
        exact_fpr = (~y & (score >= t)).sum() / (~y).sum()
        assert exact_tpr == tpr[i] and exact_fpr == fpr[i]
    # and host matrices give the same histogram
    f2, t2, _ = verification.roc_curve_binned(dn, qp, gp, bins=bins)
    assert np.array_equal(f2, fpr) and np.array_equal(t2, tpr)
    auc_b = np.trapezoid(tpr, fpr)
    auc_e = np.trapezoid(etpr, efpr)
    assert abs(auc_b - auc_e) < 1e-3


def test_topk_merge_of_gallery_parts_equals_whole_topk():
    """dali_topk_merge_f32 (the merge step of the gallery-sharded 1:N identification): per-part fused
    top-k lists stacked as an all-gather leaves them, merged by one warp per query, equal the top-k
    over the whole gallery bit for bit -- including duplicated gallery rows (ties by gallery id)
    and a part shorter than k (padding ids -1)."""
    import ctypes
    from daliid_b200 import _lib, metrics
    from daliid_b200._lib import c_vp
    g = torch.Generator().manual_seed(4)
    qf = torch.randn(500, 96, generator=g).cuda()
    gf = torch.randn(3007, 96, generator=g).cuda()
    gf[2000:2100] = gf[100:200]                    # exact ties across parts
    k = 20
    ev, ei = metrics.topk_features(qf, gf, k=k)
    bounds = [(0, 1500), (1500, 2995), (2995, 3007)]   # the last part has 12 < k rows
    vs, js = zip(*[metrics.topk_features(qf, gf[a:b].contiguous(), k=k, g_base=a) for a, b in bounds])
    gv, gi = torch.stack(vs).contiguous(), torch.stack(js).contiguous()
    assert int((gi[2] == -1).sum()) == 500 * (k - 12)
    out_v, out_i = torch.empty_like(ev), torch.empty_like(ei)
    ctx = _lib.get_ctx(0)
    ctx.attach_torch_stream()
    ctx.check(ctx.lib.dali_topk_merge_f32(ctx.h, c_vp(gv.data_ptr()), c_vp(gi.data_ptr()), 3, 500, k, 0,
                                          c_vp(out_v.data_ptr()), c_vp(out_i.data_ptr())))
    assert torch.equal(out_i, ei) and torch.equal(out_v, ev)

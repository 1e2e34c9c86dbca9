"""CPU restatement of the reference's meta-recognition score fusion.

TEST INFRASTRUCTURE ONLY -- imported by tests/ (and nothing in the product path).

Follows /root/reference/Person-ReID/evaluate.py:
    Meta_Recognition.mrfuse                   610-627   (commented call site: 277)
    Meta_Recognition.metarec(use_columns=False) 599-608
    libmr.FitHigh / _weibullFitting           429-432, 475-493
    libmr._fit (Newton MLE of the shape)      531-580
    libmr.compute_weibull_object / wscore     434-473
    torch.distributions.Weibull.cdf           (TransformedDistribution: affine^-1, power^-1,
                                               Exponential(1).cdf, then sign*(v-0.5)+0.5)

Pinned against the reference's own classes executed here: tests/golden/mrfuse.npz
(tests/golden/make_golden_mrfuse.py).  torch CPU ops are used where the reference's dtype
promotion matters (fp32 log / mean, fp64 pow).
"""
import numpy as np
import torch

TOPK = 20
TRANSLATE = 1


def weibull_fit(data, iters=100, eps=1e-6):
    """data: fp32 [N, T], every value > 0.  Returns float64 [N, 2] = (shape, scale); rows that never
    reach |dk| < eps keep (0, 0), rows whose f turned NaN keep (nan, nan) (evaluate.py:531-580)."""
    data = torch.as_tensor(data, dtype=torch.float32)
    N = data.shape[0]
    k = torch.ones(N, dtype=torch.float64)
    k_prev = k.clone()
    ln_x = torch.log(data)  # fp32
    mean_ln = torch.mean(ln_x, dim=1)  # fp32
    params = torch.zeros(N, 2, dtype=torch.float64)
    open_rows = torch.ones(N, dtype=torch.bool)
    for _ in range(iters):
        if not open_rows.any():
            break
        x_k = data ** k[:, None]  # fp64
        w = x_k * ln_x
        fg = x_k.sum(1)
        ff = w.sum(1)
        ff_prime = (w * ln_x).sum(1)
        r = ff / fg
        f = r - mean_ln - 1.0 / k
        f_prime = (ff_prime / fg - r ** 2) + 1.0 / (k * k)
        k = k - f / f_prime
        params[open_rows & torch.isnan(f)] = float("nan")
        open_rows = open_rows & ~((k - k_prev).abs() < eps)
        done = ~open_rows
        params[done, 0] = k[done]
        lam = torch.mean(data ** k[:, None], dim=1) ** (1.0 / k)
        params[done, 1] = lam[done]
        k_prev = k.clone()
    return params.numpy()


def tail_of_columns(scores, topk=TOPK):
    """The fitting data of metarec(use_columns=False): per gallery column, the Q-topk-1 largest
    scores after the per-row top-``topk`` entries were zeroed; returns (sorted fp32 [G, tail], small [G])."""
    s = torch.as_tensor(scores, dtype=torch.float32).clone()
    tval, tidx = torch.topk(s, topk, dim=1)
    s = s - torch.zeros_like(s).scatter_(1, tidx, tval)
    t = torch.nan_to_num(s.T, 0)
    tail = int(t.shape[1] - topk - 1)
    srt = torch.topk(t, tail, dim=1, largest=True, sorted=True).values
    return srt, srt[:, tail - 1]


def weibull_cdf(value, scale, shape):
    """torch.distributions.Weibull(scale, shape).cdf(value) spelled out, fp64."""
    value = torch.as_tensor(value, dtype=torch.float64)
    scale = torch.as_tensor(scale, dtype=torch.float64)
    shape = torch.as_tensor(shape, dtype=torch.float64)
    expo = 1.0 / shape
    y = (value / scale).pow(1.0 / expo)
    v = 1.0 - torch.exp(-y)
    sign = expo.sign() * scale.sign()
    return sign * (v - 0.5) + 0.5


def metarec(scores, topk=TOPK):
    """Weights [Q, G] float64, plus the fit (shape, scale) [G, 2] and the tail minimum [G]."""
    s = torch.as_tensor(scores, dtype=torch.float32)
    srt, small = tail_of_columns(s, topk)
    fit = weibull_fit(srt + TRANSLATE - small[:, None])
    d = (s + TRANSLATE - small[None, :]).clamp(min=0)
    w = weibull_cdf(d, fit[:, 1][None, :], fit[:, 0][None, :])
    return torch.nan_to_num(w, 0).numpy(), fit, small.numpy()


def mrfuse(scores01, scores02, scores03, topk=TOPK):
    """(w1*s1 + w2*s2 + w3*s3) / (w1 + w2 + w3), float64 [Q, G] (evaluate.py:610-627)."""
    s = [torch.as_tensor(x, dtype=torch.float32) for x in (scores01, scores02, scores03)]
    w = [torch.from_numpy(metarec(x, topk)[0]) for x in s]
    return ((w[0] * s[0] + w[1] * s[1] + w[2] * s[2]) / (w[0] + w[1] + w[2])).numpy()

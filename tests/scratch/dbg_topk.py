import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from daliid_b200 import metrics, _lib
g = torch.Generator().manual_seed(12)
Q, G, D = 300, 7777, 96
qf = torch.randn(Q, D, generator=g).cuda()
gf = torch.randn(G, D, generator=g).cuda()
for prec in ("tf32c", "tf32"):
    d = metrics.compute_distance_matrix(qf, gf, "cosine", prec)
    for k in (20, 64, 100, 128):
        for rep in range(2):
            v, i = metrics.topk_features(qf, gf, k=k, precision=prec)
            ev, ei = metrics.topk_identify(d, k=k)
            bad_rows = (i != ei).any(1).nonzero().flatten()
            print(prec, k, rep, "bad rows", bad_rows.numel(), "fallbacks", _lib.get_ctx(0).fallback_count())
            if bad_rows.numel():
                r = int(bad_rows[0])
                true_d = d[r, i[r].long()]
                print(" row", r, "first bad pos", int((i[r] != ei[r]).nonzero()[0]))
                print(" reported v == true d at reported idx:", bool(torch.equal(v[r], true_d)))
                wrong = (v[r] != true_d).nonzero().flatten()
                print(" #wrong values", wrong.numel(), "idx", i[r][wrong][:10].tolist(), "v", v[r][wrong][:5].tolist(), "true", true_d[wrong][:5].tolist())
                missing = set(ei[r].tolist()) - set(i[r].tolist())
                print(" missing", sorted(missing)[:20])

"""Golden vectors for the meta-recognition fusion.

Runs the reference's OWN classes -- ``libmr`` and ``Meta_Recognition`` of
/root/reference/Person-ReID/evaluate.py (394-627) -- on small seeded score matrices and stores
inputs + outputs in ``mrfuse.npz``.  The module itself cannot be imported (it imports torchreid at
the top, which is not installed), so the two class definitions are cut out of the file with ``ast``
and executed unmodified against the installed torch.  Needs /root/reference: run in the build
container only; the tests read the committed .npz.

    python tests/golden/make_golden_mrfuse.py
"""
import ast
import contextlib
import io
import os

import numpy as np
import torch

REF = "/root/reference/Person-ReID/evaluate.py"


def reference_classes():
    tree = ast.parse(open(REF).read())
    ns = {"torch": torch, "np": np}
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name in ("libmr", "Meta_Recognition"):
            exec(compile(ast.Module([node], []), REF, "exec"), ns)
    return ns["Meta_Recognition"]


def scores(seed, Q, G, D, ids, sigma):
    g = torch.Generator().manual_seed(seed)
    c = torch.randn(ids, D, generator=g)
    qp = torch.randint(0, ids, (Q,), generator=g)
    gp = torch.randint(0, ids, (G,), generator=g)
    q = c[qp] + sigma * torch.randn(Q, D, generator=g)
    ga = c[gp] + sigma * torch.randn(G, D, generator=g)
    q = q / q.norm(dim=1, keepdim=True)
    ga = ga / ga.norm(dim=1, keepdim=True)
    return q @ ga.T


def run(MR, s):
    mr = MR()
    out = {}
    with contextlib.redirect_stdout(io.StringIO()):
        fused = mr.mrfuse(*[x.clone() for x in s])
        for m, x in enumerate(s):
            w = MR().metarec(x.clone(), 20, use_columns=False)
            fit = MR()
            fit.metarec(x.clone(), 20, use_columns=False)
            out[f"w{m}"] = w.numpy()
            out[f"fit{m}"] = fit.mr.wbFits.numpy()
            out[f"small{m}"] = fit.mr.smallScoreTensor.numpy()[:, 0]
    out["fused"] = fused
    return out


def main():
    MR = reference_classes()
    blob = {}
    # case a: plain; case b: a constant (all-zero) gallery column, a NaN score, duplicated query rows
    for name, (Q, G, D) in {"a": (64, 48, 32), "b": (96, 40, 16)}.items():
        s = [scores(10 * (ord(name) - 96) + m, Q, G, D, 16, 1.5) for m in range(3)]
        if name == "b":
            # which of several equal scores torch.topk "kills" at the top-20 boundary of a row is
            # implementation defined, so ties are placed where the result does not depend on it:
            # inside gallery columns (duplicated query rows), not inside query rows
            s[0][:, 5] = 0.0
            s[1][3, 7] = float("nan")
            s[2][7, :] = s[2][8, :]
            s[2][9, :] = s[2][8, :]
        for m in range(3):
            blob[f"{name}_s{m}"] = s[m].numpy()
        for k, v in run(MR, s).items():
            blob[f"{name}_{k}"] = v
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mrfuse.npz")
    np.savez_compressed(path, **blob)
    print("wrote", path, {k: v.shape for k, v in blob.items() if k.endswith("fused")})


if __name__ == "__main__":
    main()

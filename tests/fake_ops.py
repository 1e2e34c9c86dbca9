"""TEST-ONLY stand-in for daliid_b200.sharded.CudaOps so the multi-rank host logic
(slab partition, the two all-reduces, the top-k all_gather + merge) can run under gloo on
CPU.  It is numpy, lives under tests/, and is never importable from the product package."""
import numpy as np
import torch


def dist_key(d):
    d = np.asarray(d, dtype=np.float32) + np.float32(0.0)
    b = d.view(np.uint32)
    k = np.where(b & np.uint32(0x80000000), ~b, b | np.uint32(0x80000000))
    return np.where(np.isnan(d), np.uint32(0xFFFFFFFE), k).astype(np.uint32)


class FakeOps:
    def distmat(self, qf, gf_slab, metric, precision, normalize):
        assert metric == "cosine"
        q = qf / torch.norm(qf, dim=1, keepdim=True)
        g = gf_slab / torch.norm(gf_slab, dim=1, keepdim=True)
        return 1.0 - torch.mm(q, g.T)

    def plan(self, qp, gp, qc, gc):
        off, gid, nv = [0], [], []
        for q in range(len(qp)):
            m = np.flatnonzero(gp == qp[q])
            valid = m[gc[m] != qc[q]]
            junk = m[gc[m] == qc[q]]
            gid += list(valid) + list(junk)
            nv.append(len(valid))
            off.append(len(gid))
        return dict(off=np.array(off), gid=np.array(gid, dtype=np.int64), nv=np.array(nv), G=len(gp))

    def plan_destroy(self, plan):
        pass

    def gather_keys(self, plan, dist_slab, g0):
        d = dist_slab.numpy()
        keys = np.zeros(max(len(plan["gid"]), 1), dtype=np.uint32)
        for q in range(len(plan["nv"])):
            for t in range(plan["off"][q], plan["off"][q + 1]):
                loc = plan["gid"][t] - g0
                if 0 <= loc < d.shape[1]:
                    keys[t] = dist_key(d[q, loc])
        return torch.from_numpy(keys.view(np.int32).copy())

    def count(self, plan, dist_slab, g0, keys):
        d = dist_slab.numpy()
        k = keys.numpy().view(np.uint32).astype(np.uint64)
        counts = np.zeros_like(keys.numpy())
        cols = np.arange(d.shape[1], dtype=np.uint64) + np.uint64(g0)
        for q in range(len(plan["nv"])):
            comp = (dist_key(d[q]).astype(np.uint64) << np.uint64(32)) | cols
            for t in range(plan["off"][q], plan["off"][q] + plan["nv"][q]):
                thr = (k[t] << np.uint64(32)) | np.uint64(plan["gid"][t])
                counts[t] = int((comp < thr).sum())
        return torch.from_numpy(counts)

    def finalize(self, plan, keys, counts, Q, G, max_rank, accum):
        k = keys.numpy().view(np.uint32).astype(np.uint64)
        c = counts.numpy()
        f32 = np.float32
        ap = np.full(Q, np.nan)
        first = np.full(Q, -1, dtype=np.int32)
        cmc_cnt = np.zeros(max_rank, dtype=np.int64)
        for q in range(Q):
            o, nv, e = plan["off"][q], plan["nv"][q], plan["off"][q + 1]
            if nv == 0:
                continue
            comp = (k[o:e] << np.uint64(32)) | plan["gid"][o:e].astype(np.uint64)
            ranks = sorted(int(c[o + t]) - int((comp[nv:] < comp[t]).sum()) + 1 for t in range(nv))
            first[q] = ranks[0]
            if ranks[0] <= max_rank:
                cmc_cnt[ranks[0] - 1:] += 1
            s = f32(0)
            for i, r in enumerate(ranks, 1):
                s = f32(float(s) + float(i) / float(r))
            ap[q] = f32(s / f32(nv))
        nvq = int((first > 0).sum())
        assert nvq > 0, "Error: all query identities do not appear in gallery"
        m = f32(0)
        for q in range(Q):
            if first[q] > 0:
                m = f32(m + f32(ap[q]))
        return ((cmc_cnt.astype(np.float32) / f32(nvq)).astype(np.float32), float(f32(m / f32(nvq))),
                {"ap": ap, "first_rank": first, "num_valid": nvq})

    def topk(self, distmat, k, largest, col_ids=None):
        d = np.asarray(distmat, dtype=np.float32)
        ids = np.asarray(col_ids) if col_ids is not None else np.tile(np.arange(d.shape[1]), (d.shape[0], 1))
        key = dist_key(d).astype(np.uint64)
        if largest:
            key = (~key.astype(np.uint32)).astype(np.uint64)
        comp = (key << np.uint64(32)) | ids.astype(np.uint32).astype(np.uint64)
        order = np.argsort(comp, axis=1, kind="stable")[:, :k]
        return (torch.from_numpy(np.take_along_axis(d, order, 1)),
                torch.from_numpy(np.take_along_axis(ids, order, 1).astype(np.int32)))

    def topk_features(self, qf, gf_slab, k, metric, precision, normalize, largest, g_base):
        d = self.distmat(qf, gf_slab, metric, precision, normalize).numpy()
        v, i = self.topk(d, min(k, d.shape[1]), largest)
        if d.shape[1] < k:
            pad = k - d.shape[1]
            v = torch.cat([v, torch.full((d.shape[0], pad), float("inf"))], 1)
            i = torch.cat([i, torch.full((d.shape[0], pad), -1 - g_base, dtype=torch.int32)], 1)
        return v, i + g_base

"""Gallery-sharded evaluation: one process per GPU (``torch.distributed``, NCCL over NVLink).

The reference has no multi-GPU evaluation (its hot path is single-process CPU code); this
is the scaling axis the build adds.  Rank ``r`` holds ALL queries and a
contiguous gallery slab ``[g0, g0+Gs)``; slabs are contiguous in gallery index so the global
tie-break ``(distance, gallery index)`` is preserved and the result is bit-identical to the
single-GPU one.  The only exchange steps are two tiny all-reduces:

    slab distmat -> gather match keys -> all_reduce(SUM) -> count below -> all_reduce(SUM)
                 -> CMC/mAP epilogue (every rank, identical result)

and, for top-k identification, an all_gather of the per-slab top-k followed by a k-way merge.
Everything heavy (contraction, counting, selection) is a local kernel of the C-ABI library.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch
import torch.distributed as dist

from . import _lib, metrics
from ._lib import ACCUMS, c_vp, get_ctx, p_i32


def slab_bounds(G: int, world: int, rank: int):
    """Contiguous, balanced partition of the gallery: returns (g0, Gs)."""
    base, rem = divmod(G, world)
    g0 = rank * base + min(rank, rem)
    return g0, base + (1 if rank < rem else 0)


class CudaOps:
    """Local building blocks backed by the C-ABI (``dali_rank_*``, ``dali_distmat_f32`` ...)."""

    def __init__(self, device=None):
        self.ctx = get_ctx(device)
        self.device = torch.device(f"cuda:{self.ctx.device}")

    def distmat(self, qf, gf_slab, metric, precision, normalize):
        # host features are streamed in (chunked H2D overlapped with compute); the slab
        # matrix always stays on this rank's GPU
        return metrics.compute_distance_matrix(qf, gf_slab, metric, precision, normalize,
                                               device=self.ctx.device, padded=True)

    def plan(self, q_pid, g_pid, q_cam, g_cam):
        self.ctx.attach_torch_stream()
        h = c_vp()
        self._labels = (q_pid, g_pid, q_cam, g_cam)
        self.ctx.check(self.ctx.lib.dali_rank_plan_create(
            self.ctx.h, p_i32(q_pid), p_i32(g_pid), p_i32(q_cam), p_i32(g_cam),
            len(q_pid), len(g_pid), ctypes.byref(h)))
        return h

    def plan_destroy(self, plan):
        self.ctx.lib.dali_rank_plan_destroy(plan)

    def num_matches(self, plan):
        return int(self.ctx.lib.dali_rank_plan_num_matches(plan))

    def gather_keys(self, plan, dist_slab, g0, out_ptr=None):
        """Keys of the matches whose gallery item lies in this slab (0 elsewhere).  ``out_ptr``:
        write into that device buffer (a peer-exchange buffer) instead of a fresh tensor."""
        self.ctx.attach_torch_stream()
        M = self.num_matches(plan)
        keys = None
        if out_ptr is None:
            keys = torch.zeros(max(M, 1), dtype=torch.int32, device=self.device)
            out_ptr = keys.data_ptr()
        Q, Gs = dist_slab.shape
        self.ctx.check(self.ctx.lib.dali_rank_gather_keys(
            self.ctx.h, plan, c_vp(dist_slab.data_ptr()), max(dist_slab.stride(0), Gs, 1) if Q else max(Gs, 1),
            g0, Gs, c_vp(out_ptr)))
        return keys

    def count(self, plan, dist_slab, g0, keys, out_ptr=None):
        self.ctx.attach_torch_stream()
        counts = None
        if out_ptr is None:
            counts = torch.zeros_like(keys)
            out_ptr = counts.data_ptr()
        Q, Gs = dist_slab.shape
        self.ctx.check(self.ctx.lib.dali_rank_count(
            self.ctx.h, plan, c_vp(dist_slab.data_ptr()), max(dist_slab.stride(0), Gs, 1) if Q else max(Gs, 1),
            g0, Gs, c_vp(keys.data_ptr()), c_vp(out_ptr)))
        return counts

    def finalize(self, plan, keys, counts, Q, G, max_rank, accum):
        self.ctx.attach_torch_stream()
        max_rank = min(max_rank, G)
        cmc = np.zeros(max_rank, dtype=np.float32)
        mAP = ctypes.c_double(0.0)
        ap = np.zeros(Q, dtype=np.float64)
        first = np.zeros(Q, dtype=np.int32)
        nvalid = ctypes.c_int64(0)
        self.ctx.check(self.ctx.lib.dali_rank_finalize(
            self.ctx.h, plan, c_vp(keys.data_ptr()), c_vp(counts.data_ptr()), max_rank,
            ACCUMS[accum], cmc.ctypes.data_as(_lib.c_f32p), ctypes.byref(mAP),
            ap.ctypes.data_as(_lib.c_f64p), first.ctypes.data_as(_lib.c_i32p), ctypes.byref(nvalid)))
        return cmc, float(mAP.value), {"ap": ap, "first_rank": first, "num_valid": int(nvalid.value)}

    def topk(self, distmat, k, largest, col_ids=None):
        return metrics.topk_identify(distmat, k, largest, col_ids)

    def topk_features(self, qf, gf_slab, k, metric, precision, normalize, largest, g_base):
        return metrics.topk_features(qf, gf_slab, k, metric, precision, normalize, largest, g_base)


def _world(group):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def _all_reduce_sum(t, group):
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


class PeerExchange:
    """The two exchange steps over NVLink peer memory instead of NCCL (``csrc/peer.cu``).

    Every rank allocates one block (two int32 buffers + flag words), the CUDA IPC handles are
    exchanged once through ``torch.distributed`` (the plumbing), and from then on an exchange is ONE
    kernel per rank: signal "my contribution is complete" into every peer's flag word, wait for
    all peers, sum the peers' buffers with plain loads.  ``gather_keys`` / ``count`` write their
    contribution straight into the block.  Collective: all ranks of the group make the same calls
    in the same order.  Single node, world <= 8, CUDA only."""

    def __init__(self, ops, capacity, group=None):
        self.ops = ops
        self.group = group
        self.world, self.rank = _world(group)
        self.h = c_vp()
        ctx = ops.ctx
        ctx.check(ctx.lib.dali_peer_create(ctx.h, self.rank, self.world, int(capacity), ctypes.byref(self.h)))
        self.capacity = int(ctx.lib.dali_peer_capacity(self.h))
        mine = ctypes.create_string_buffer(64)
        ctx.check(ctx.lib.dali_peer_ipc_handle(self.h, mine))
        if self.world > 1:
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(mine.raw), group=group)
            blob = ctypes.create_string_buffer(b"".join(handles), 64 * self.world)
            err = None
            try:
                ctx.check(ctx.lib.dali_peer_connect(self.h, blob))
            except _lib.DaliError as e:  # keep the collective sequence identical on every rank
                err = e
            dist.barrier(group=group)  # every rank has mapped every block before anyone signals
            if err is not None:
                ctx.lib.dali_peer_destroy(self.h)
                self.h = c_vp()
                raise err
        self.out = [torch.empty(self.capacity, dtype=torch.int32, device=ops.device) for _ in range(2)]

    def buffer_ptr(self, which):
        return self.ops.ctx.lib.dali_peer_buffer(self.h, which)

    def allreduce(self, which, n):
        """Sum over ranks of buffer ``which`` -> a device tensor of ``n`` int32."""
        ctx = self.ops.ctx
        ctx.attach_torch_stream()
        out = self.out[which]
        ctx.check(ctx.lib.dali_peer_allreduce_i32(ctx.h, self.h, which, c_vp(out.data_ptr()), int(n)))
        return out[:max(int(n), 1)]

    def close(self):
        if self.h:
            if self.world > 1:
                torch.cuda.synchronize(self.ops.device)
                dist.barrier(group=self.group)  # nobody unmaps while a peer may still read
            self.ops.ctx.lib.dali_peer_destroy(self.h)
            self.h = c_vp()


_peer_cache = {}


_peer_disabled = set()


def peer_exchange(ops, need, group=None):
    """Process-wide PeerExchange for this (device, group); re-created collectively when a larger
    capacity is needed (``need`` = number of matches, identical on every rank).  Returns None --
    on every rank -- when peer mapping is not possible (CUDA IPC refused on some rank): the
    caller then uses the NCCL exchange."""
    key = (ops.ctx.device, id(group))
    if key in _peer_disabled:
        return None
    px = _peer_cache.get(key)
    if px is None or px.capacity < need:
        if px is not None:
            px.close()
            _peer_cache.pop(key, None)
        ok = 1
        try:
            px = PeerExchange(ops, max(int(need * 1.25), 1 << 16), group)
        except _lib.DaliError:
            px, ok = None, 0
        world, _ = _world(group)
        if world > 1:  # all ranks take the same branch
            t = torch.tensor([ok], dtype=torch.int32, device=ops.device)
            dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
            ok = int(t.item())
        if not ok:
            if px is not None:
                px.close()
            _peer_disabled.add(key)
            return None
        _peer_cache[key] = px
    return px


def share_queries(qf, device, group=None):
    """Queries are replicated on every rank, so when they live in HOST memory each rank uploads only
    its 1/world share over PCIe and the ranks exchange the shares over NVLink (one all_gather):
    per-rank host traffic drops from Q*D + Gs*D to Q*D/world + Gs*D floats -- 16 % less at the
    Market shape on 8 GPUs, where the ranks compete for the host's memory path.  Every rank must
    pass the SAME query matrix (the sharded evaluation requires that anyway).  Returns a ``[Q, D]``
    fp32 CUDA tensor; device inputs and single-rank runs are returned unchanged."""
    world, rank = _world(group)
    if world == 1 or getattr(qf, "is_cuda", False) or dist.get_backend(group) != "nccl":
        return qf
    t = qf if isinstance(qf, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(qf, dtype=np.float32))
    if t.dim() != 2 or t.dtype != torch.float32 or not t.is_contiguous():
        return qf
    Q, D = t.shape
    chunk = (Q + world - 1) // world
    dev = torch.device(f"cuda:{device}") if not isinstance(device, torch.device) else device
    full = torch.empty((world * chunk, D), dtype=torch.float32, device=dev)
    mine = torch.zeros((chunk, D), dtype=torch.float32, device=dev) if (rank + 1) * chunk > Q else \
        torch.empty((chunk, D), dtype=torch.float32, device=dev)
    r0, r1 = rank * chunk, min(Q, (rank + 1) * chunk)
    if r1 > r0:
        mine[:r1 - r0].copy_(t[r0:r1], non_blocking=True)
    dist.all_gather_into_tensor(full, mine, group=group)
    return full[:Q]


def gather_gallery_labels(g_pid_slab, g_cam_slab, group=None):
    """All ranks learn the labels of the whole gallery (two int32 vectors; tiny) and the
    slab offsets.  Returns (g_pid_all, g_cam_all, g0_of_this_rank, sizes)."""
    world, rank = _world(group)
    pid = np.ascontiguousarray(g_pid_slab, dtype=np.int32)
    cam = np.ascontiguousarray(g_cam_slab, dtype=np.int32)
    if world == 1:
        return pid, cam, 0, [len(pid)]
    objs = [None] * world
    dist.all_gather_object(objs, (pid, cam), group=group)
    sizes = [len(o[0]) for o in objs]
    return (np.concatenate([o[0] for o in objs]), np.concatenate([o[1] for o in objs]),
            int(sum(sizes[:rank])), sizes)


def evaluate_rank_sharded(dist_slab, g0, q_pids, g_pids_all, q_camids, g_camids_all, max_rank=50,
                          accum="cy_f32", group=None, ops=None, return_details=False, exchange="auto"):
    """CMC/mAP from a per-rank distance slab ``[Q, Gs]`` (columns = gallery ``g0 .. g0+Gs``).
    Labels are those of the whole gallery.  Every rank returns the same ``(cmc, mAP)``.

    ``dist_slab`` may still be in flight on the GPU (stream-ordered): the host-side plan
    construction below overlaps the contraction that produces it."""
    ops = ops or CudaOps(dist_slab.device.index if dist_slab.is_cuda else None)
    qp, gp = metrics.canonicalize_labels(q_pids, g_pids_all)
    qc, gc = metrics.canonicalize_labels(q_camids, g_camids_all)
    Q, G = len(qp), len(gp)
    if dist_slab.shape[0] != Q or g0 < 0 or g0 + dist_slab.shape[1] > G:
        raise ValueError("slab does not fit the label arrays")
    if G < max_rank:
        max_rank = G
        print("Note: number of gallery samples is quite small, got {}".format(G))
    world, _ = _world(group)
    use_peer = (exchange == "peer" or (exchange == "auto" and world > 1 and isinstance(ops, CudaOps)
                                       and dist.get_backend(group) == "nccl"))
    plan = ops.plan(qp, gp, qc, gc)
    try:
        px = None
        if use_peer:
            M = ops.num_matches(plan)
            px = peer_exchange(ops, M, group)
            if px is None and exchange == "peer":
                raise _lib.DaliError("peer exchange requested but CUDA IPC mapping failed on some rank")
        if px is not None:
            # exchange steps as single kernels over NVLink peer memory (PeerExchange)
            ops.gather_keys(plan, dist_slab, g0, out_ptr=px.buffer_ptr(0))
            keys = px.allreduce(0, M)
            ops.count(plan, dist_slab, g0, keys, out_ptr=px.buffer_ptr(1))
            counts = px.allreduce(1, M)
        else:
            keys = _all_reduce_sum(ops.gather_keys(plan, dist_slab, g0), group)
            counts = _all_reduce_sum(ops.count(plan, dist_slab, g0, keys), group)
        cmc, mAP, details = ops.finalize(plan, keys, counts, Q, G, max_rank, accum)
    finally:
        ops.plan_destroy(plan)
    return (cmc, mAP, details) if return_details else (cmc, mAP)


def evaluate_features_sharded(qf, gf_slab, g0, q_pids, g_pids_all, q_camids, g_camids_all,
                              metric="cosine", precision=metrics.DEFAULT_PRECISION,
                              normalize=None, max_rank=50, accum="cy_f32", group=None, ops=None,
                              return_details=False, exchange="auto", share_host_queries=True):
    """Features in (all queries + this rank's gallery slab), ``(cmc, mAP)`` out.  Host-resident
    queries are uploaded once in total (``share_queries``), not once per rank."""
    ops = ops or CudaOps(qf.device.index if getattr(qf, "is_cuda", False) else None)
    if normalize is None:
        normalize = metric == "cosine"
    world, _ = _world(group)
    if share_host_queries and isinstance(ops, CudaOps) and world > 1:
        qf = share_queries(qf, ops.device, group)
    if (isinstance(ops, CudaOps) and world > 1 and exchange in ("auto", "peer")
            and dist.get_backend(group) == "nccl"):
        res = _evaluate_features_one_call(ops, qf, gf_slab, g0, q_pids, g_pids_all, q_camids, g_camids_all,
                                          metric, precision, normalize, max_rank, accum, group,
                                          strict=exchange == "peer")
        if res is not None:
            return res if return_details else res[:2]
    dist_slab = ops.distmat(qf, gf_slab, metric, precision, normalize)
    if not isinstance(dist_slab, torch.Tensor):
        dist_slab = torch.from_numpy(dist_slab)
    return evaluate_rank_sharded(dist_slab, g0, q_pids, g_pids_all, q_camids, g_camids_all,
                                 max_rank, accum, group, ops, return_details, exchange)


def _evaluate_features_one_call(ops, qf, gf_slab, g0, q_pids, g_pids_all, q_camids, g_camids_all, metric,
                                precision, normalize, max_rank, accum, group, strict):
    """The whole per-rank sequence inside ``dali_eval_features_sharded_f32`` (exchange over NVLink
    peer memory): one library call per evaluation.  Returns None -- on every rank -- when peer
    mapping is unavailable, so that the caller falls back to the step-by-step NCCL path."""
    from ._lib import as_matrix
    a = as_matrix(qf, np.float32, "qf")
    b = as_matrix(gf_slab, np.float32, "gf_slab")
    if a.shape[1] != b.shape[1] or a.ld != a.shape[1] or b.ld != b.shape[1]:
        raise ValueError("feature matrices must be contiguous with equal dimension")
    qp, gp = metrics.canonicalize_labels(q_pids, g_pids_all)
    qc, gc = metrics.canonicalize_labels(q_camids, g_camids_all)
    Q, D = a.shape
    Gs, G = b.shape[0], len(gp)
    if len(qp) != Q or len(qc) != Q or len(gc) != G or g0 < 0 or g0 + Gs > G:
        raise ValueError("label arrays do not match the feature matrices / slab")
    if G < max_rank:
        max_rank = G
        print("Note: number of gallery samples is quite small, got {}".format(G))
    ctx = ops.ctx
    ctx.attach_torch_stream()
    need = getattr(ops, "_last_matches", 1 << 16)
    for _ in range(3):
        px = peer_exchange(ops, need, group)
        if px is None:
            if strict:
                raise _lib.DaliError("peer exchange requested but CUDA IPC mapping failed on some rank")
            return None
        cmc = np.zeros(max_rank, dtype=np.float32)
        mAP = ctypes.c_double(0.0)
        ap = np.zeros(Q, dtype=np.float64)
        first = np.zeros(Q, dtype=np.int32)
        nvalid = ctypes.c_int64(0)
        matches = ctypes.c_int64(0)
        rc = ctx.lib.dali_eval_features_sharded_f32(
            ctx.h, px.h, c_vp(a.ptr), Q, c_vp(b.ptr), Gs, D, int(g0), G, p_i32(qp), p_i32(gp), p_i32(qc),
            p_i32(gc), metrics._enum(metrics.METRICS, metric, "metric"), metrics._precision(precision, normalize, D),
            1 if normalize else 0, int(max_rank), ACCUMS[accum], cmc.ctypes.data_as(_lib.c_f32p),
            ctypes.byref(mAP), ap.ctypes.data_as(_lib.c_f64p), first.ctypes.data_as(_lib.c_i32p),
            ctypes.byref(nvalid), ctypes.byref(matches))
        ops._last_matches = int(matches.value)
        if rc == _lib.ERR_PEER_CAPACITY:  # same answer on every rank: grow the block collectively
            need = int(matches.value)
            continue
        ctx.check(rc)
        return cmc, float(mAP.value), {"ap": ap, "first_rank": first, "num_valid": int(nvalid.value)}
    raise _lib.DaliError("peer block could not be sized for the number of matches")


def topk_features_sharded(qf, gf_slab, g0, k=20, metric="cosine", precision=metrics.DEFAULT_PRECISION,
                          normalize=None, largest=False, group=None, ops=None):
    """1:N identification over a sharded gallery: per-slab fused distance + top-k, all_gather of
    the ``[Q,k]`` candidates, k-way merge with the global (value, gallery id) order."""
    ops = ops or CudaOps(qf.device.index if getattr(qf, "is_cuda", False) else None)
    if normalize is None:
        normalize = metric == "cosine"
    if isinstance(ops, CudaOps):
        qf = share_queries(qf, ops.device, group)
    vals, ids = ops.topk_features(qf, gf_slab, k, metric, precision, normalize, largest, g0)
    world, _ = _world(group)
    if world == 1:
        return vals, ids
    vals = torch.as_tensor(vals).contiguous()
    ids = torch.as_tensor(ids).contiguous()
    if isinstance(ops, CudaOps) and k <= 32 and vals.is_cuda:
        # one all-gather per array straight into [world, Q, k], merged by one warp per query
        # (dali_topk_merge_f32): no concatenation copies, no one-CTA-per-row selection over 8 k columns
        Q = vals.shape[0]
        gv = torch.empty((world, Q, k), dtype=vals.dtype, device=vals.device)
        gi = torch.empty((world, Q, k), dtype=ids.dtype, device=ids.device)
        dist.all_gather_into_tensor(gv, vals, group=group)
        dist.all_gather_into_tensor(gi, ids, group=group)
        out_v = torch.empty_like(vals)
        out_i = torch.empty_like(ids)
        ctx = ops.ctx
        ctx.attach_torch_stream()
        ctx.check(ctx.lib.dali_topk_merge_f32(ctx.h, c_vp(gv.data_ptr()), c_vp(gi.data_ptr()), world, Q, int(k),
                                              1 if largest else 0, c_vp(out_v.data_ptr()), c_vp(out_i.data_ptr())))
        return out_v, out_i
    vs = [torch.empty_like(vals) for _ in range(world)]
    js = [torch.empty_like(ids) for _ in range(world)]
    dist.all_gather(vs, vals, group=group)
    dist.all_gather(js, ids, group=group)
    cand_v = torch.cat(vs, dim=1).contiguous()
    cand_i = torch.cat(js, dim=1).contiguous()
    # padded entries (id -1, value +-inf) lose every comparison against real candidates
    return ops.topk(cand_v, k, largest, cand_i)

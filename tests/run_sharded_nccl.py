"""Multi-GPU correctness check (run under torchrun on a box with >= 2 GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29533 tests/run_sharded_nccl.py

Every rank evaluates its gallery slab; the sharded CMC/mAP must equal, bit for bit, the
single-GPU evaluation of the whole gallery computed on each rank, and the sharded top-k must
equal the single-GPU top-k."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from daliid_b200 import metrics, sharded, synth
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    for name in ("small", "market_vit"):
        qf, gf, qp, gp, qc, gc = synth.make_config(name, device=f"cuda:{local}")
        g0, gs = sharded.slab_bounds(gf.shape[0], world, rank)
        for precision, exchange in (("auto", "peer"), ("auto", "nccl"), ("tf32c", "peer"), ("fp32", "auto")):
            cmc, mAP, det = sharded.evaluate_features_sharded(
                qf, gf[g0:g0 + gs].contiguous(), g0, qp, gp, qc, gc, precision=precision,
                return_details=True, exchange=exchange)
            e_cmc, e_map, e_det = metrics.evaluate_features(qf, gf, qp, gp, qc, gc, precision=precision,
                                                            return_details=True)
            assert np.array_equal(cmc, e_cmc) and mAP == e_map, (name, precision, exchange, mAP, e_map)
            assert np.array_equal(det["first_rank"], e_det["first_rank"])
        # host-resident features: every rank uploads 1/world of the queries, the shares travel over NVLink
        qh, gh = qf.cpu().pin_memory(), gf[g0:g0 + gs].cpu().pin_memory()
        cmc, mAP, det = sharded.evaluate_features_sharded(qh, gh, g0, qp, gp, qc, gc, return_details=True)
        e_cmc, e_map, e_det = metrics.evaluate_features(qf, gf, qp, gp, qc, gc, return_details=True)
        assert np.array_equal(cmc, e_cmc) and mAP == e_map, (name, "host features", mAP, e_map)
        assert np.array_equal(det["first_rank"], e_det["first_rank"])
        assert torch.equal(sharded.share_queries(qh, local), qf)
        v, i = sharded.topk_features_sharded(qf, gf[g0:g0 + gs].contiguous(), g0, k=20)
        ev, ei = metrics.topk_features(qf, gf, k=20)
        assert torch.equal(i, ei) and torch.equal(v, ev), name
        if rank == 0:
            print(f"{name}: sharded x{world} == single GPU (mAP {mAP:.6f})", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

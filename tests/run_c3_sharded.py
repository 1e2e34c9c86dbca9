"""BASELINE config 2 under torchrun: DeepChange-shaped evaluation (17527 x 62956, D=768) with the
gallery sharded over the ranks (strong scaling: the total work is fixed).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29551 tests/run_c3_sharded.py

Rank 0 prints one JSON line: ms per evaluation (device events, max over ranks), pairs/s, mAP."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from daliid_b200 import metrics, sharded, synth
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    qf, gf, qp, gp, qc, gc = synth.make_config("deepchange", device=dev)
    Q, G = qf.shape[0], gf.shape[0]
    g0, gs = sharded.slab_bounds(G, world, rank)
    slab = gf[g0:g0 + gs].contiguous()
    del gf

    def step():
        if world == 1:
            return metrics.evaluate_features(qf, slab, qp, gp, qc, gc)
        return sharded.evaluate_features_sharded(qf, slab, g0, qp, gp, qc, gc)

    for _ in range(3):
        cmc, mAP = step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n):
        cmc, mAP = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if rank == 0:
        print(json.dumps({"workload": f"C3 DeepChange shape: Q={Q} x G={G} (sharded x{world}) x D=768, f16x3",
                          "n_gpus": world, "ms_per_eval": ms, "pairs_per_s": Q * G / (ms * 1e-3),
                          "scaling": "strong", "mAP": mAP, "rank1": float(cmc[0])}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

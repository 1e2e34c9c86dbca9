"""BASELINE config 5 (face-ID 1:N scale) under torchrun: 100k queries, 125k gallery rows per
rank (1M over 8 ranks), D=512, fused distance + top-20 per slab, all_gather + k-way merge.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29541 tests/run_c5_sharded.py [Q] [G_per_rank] [precision]

Prints one JSON line from rank 0: time per evaluation (device events, max over ranks),
global pairs/s and the algorithmic TFLOP/s per GPU."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from daliid_b200 import sharded
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", rank))
    Q = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    Gs = int(sys.argv[2]) if len(sys.argv) > 2 else 125000
    prec = sys.argv[3] if len(sys.argv) > 3 else "auto"
    D, k = 512, 20
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    gq = torch.Generator(device=dev).manual_seed(12)
    qf = torch.randn(Q, D, generator=gq, device=dev)           # same queries on every rank
    gg = torch.Generator(device=dev).manual_seed(1000 + rank)
    gf = torch.randn(Gs, D, generator=gg, device=dev)          # this rank's slab
    g0 = rank * Gs

    def step():
        return sharded.topk_features_sharded(qf, gf, g0, k=k, precision=prec)

    for _ in range(2):
        v, i = step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 3
    e0.record()
    for _ in range(n):
        v, i = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    assert int(i.min()) >= 0 and int(i.max()) < Gs * world
    if rank == 0:
        print(json.dumps({"workload": f"C5 slab: Q={Q} x G={Gs}/GPU x D={D}, top-{k}, precision={prec}",
                          "n_gpus": world, "ms_per_eval": ms,
                          "pairs_per_s": Q * Gs * world / (ms * 1e-3),
                          "tflops_per_gpu": 2.0 * Q * Gs * D / (ms * 1e-3) / 1e12}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

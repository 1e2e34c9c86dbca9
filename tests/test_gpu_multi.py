"""Multi-GPU parity under pytest (the driver's `pytest -m gpu`): torchrun over every visible GPU,
sharded evaluation == single-GPU evaluation bit for bit, with the NVLink peer exchange and with
NCCL.  Skips on a one-GPU box (the gloo tests in test_sharded_gloo.py cover the host logic)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(script, nproc, port):
    env = dict(os.environ)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", script)]
    p = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-3000:] + "\n" + p.stderr[-3000:]
    return p.stdout


def test_sharded_evaluation_equals_single_gpu_on_all_gpus():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    out = _run("run_sharded_nccl.py", min(n, 8), 29541)
    assert "market_vit: sharded" in out


def test_two_devices_in_one_process():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least two GPUs")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "run_two_devices.py")], cwd=ROOT,
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]

"""The warp-per-query form of the counting kernel (``rank_count_wpq_kernel``, opt-in through
``DALI_RANK_WPQ``; the switch is read once per process) must give the results of the default kernel:
the ranking tests are run again in a child process with the switch set so that every query with at
most 62 matches takes it, whatever the number of queries."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_rank_tests_pass_with_the_warp_per_query_kernel():
    env = dict(os.environ, DALI_RANK_WPQ="2")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_rank.py"),
                        os.path.join(ROOT, "tests", "test_gpu_fused_count.py"), "-x", "-q", "-m", "gpu",
                        "-p", "no:cacheprovider"],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]

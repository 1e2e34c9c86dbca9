"""bench.py prints ONE JSON line with the keys the driver reads.  The reference arm (CPU) is checked
here on the CPU; our arm needs a B200 (gpu marker)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
             "scaling", "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"}


def _run(args, env=None):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, env=env,
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_contract():
    d = _run(["--impl", "reference", "--steps", "1", "--warmup", "0"])
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"].startswith("QxG pairs/s") and d["unit"] == "pairs/s" and d["higher_is_better"] is True
    assert "workload" in d["config"] and d["value"] > 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] and "sample" in d["cpu_baseline"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_do_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "0"], cwd=ROOT, env=env, capture_output=True, text=True,
                         timeout=300)
    assert out.returncode == 0 and not [l for l in out.stdout.splitlines() if l.startswith("{")]


@pytest.mark.gpu
def test_our_arm_contract():
    d = _run(["--steps", "3", "--warmup", "3", "--no-c5", "--no-modes"])
    assert BASE_KEYS <= set(d) and "impl" not in d
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] >= 3 and d["scaling"] == "weak"
    assert d["value"] > 1e9 and d["ms_per_step"] > 0 and d["data"].startswith("synthetic")
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    e = d["e2e"]
    assert e["unit"] == d["unit"] and 0 < e["value"] < d["value"]
    assert e["h2d_bytes_per_step"] > 1e8 and e["d2h_bytes_per_step"] > 0
    assert d["gpu_launches"] >= 3 * 3
    r = d["roofline"]
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and 0 < r["frac"] < 1
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["traffic"] is None or r["traffic"] > 1e8      # null unless measured in the run (never a constant)
    assert d["e2e_pageable"]["value"] > 0 and d["h2d_ceiling"]["gbs_aggregate"] > 1
    assert d["c1_market_vit"]["under_target"] is True and d["c1_market_vit"]["ms_per_step"] < 50
    assert d["c3_deepchange"]["scaling"] == "strong" and d["c4_fusion3"]["ms_per_step"] > 0
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0 and c["unit"] == d["unit"]
    assert d["value"] > 50 * c["value"]

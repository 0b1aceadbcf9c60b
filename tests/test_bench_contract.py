"""bench.py prints exactly ONE JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"}


def _run(args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, stdout=subprocess.PIPE,
                         stderr=subprocess.PIPE, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _run(["--impl", "reference", "--steps", "2", "--warmup", "0", "--region", "64"])
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"].startswith("chunks/sec gen+mesh") and d["unit"] == "chunks/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "chunks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["scaling"] == "strong" and d["detail"]["chunks_per_step"] == 4 * 64
    # both arms print the same `config` for the same region
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.workload_config(64)


@pytest.mark.gpu
def test_own_arm_line():
    # a 256 x 256 region (8 batches of 8192 chunks) keeps the test short; the default is configs[2]'s 1024 x 1024
    d = _run(["--steps", "2", "--warmup", "3", "--region", "256"])
    assert BASE_KEYS | {"roofline", "clocks", "gpu_launches", "quads_per_s", "parity"} <= set(d)
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["dtype"] == "f32" and d["data"] == "synthetic"
    assert d["scaling"] == "strong" and d["detail"]["chunks_per_step"] == 256 * 256
    assert d["value"] > 1e5 and d["e2e"]["value"] > 1e4 and d["e2e"]["value"] < d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 1e6
    assert 0 < d["e2e"]["link_frac"] < 1.1 and d["e2e"]["link_peak_gbs"] > 10
    assert d["gpu_launches"] >= 2 * 8 * 4
    par = d["parity"]
    assert par["region_digest"]["chunks"] == 256 * 256 and par["quads_match_e2e_pass"] is True
    assert par["oracle_sample"]["pass_counts_equal_oracle"] is True
    assert par["oracle_sample"]["instance_bytes_equal_oracle"] is True
    assert d["ring32x32"]["two_in_flight"]["chunks_per_s"] > 1e5
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and 0 < r["frac"] < 1.2
    assert d["roofline_mesh"]["bound"] == "hbm" and 0 < d["roofline_mesh"]["frac"] < 1.2
    assert d["cpu_baseline"]["value"] > 0
    assert d["roofline_mesh"]["steady_state"]["frac"] > 0
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert 1e4 < d["e2e_pageable_sample"]["value"] < d["value"] and d["e2e_pageable_sample"]["d2h_gbs"] > 1

"""bench.py prints exactly one JSON line with the keys the driver reads (both arms)."""
import json
import os
import subprocess
import sys

import pytest

import helpers as H

BENCH = os.path.join(H.ROOT, "bench.py")
COMMON = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
          "scaling", "vs_baseline", "dtype", "data", "config", "e2e"]


def _one_json_line(cmd):
    out = subprocess.run([sys.executable, BENCH] + cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, f"stdout must be one JSON line, got {len(lines)}: {out.stdout[:400]}"
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _one_json_line(["--impl", "reference", "--steps", "1", "--warmup", "0", "--size", "64"])
    for k in COMMON + ["impl", "cpu_baseline"]:
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "7-pt diffusion GLUP/s" and d["unit"] == "GLUP/s"
    assert d["value"] > 0 and d["higher_is_better"] is True
    assert d["e2e"] == {"value": d["value"], "unit": "GLUP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]


@pytest.mark.gpu
def test_b200_arm_line():
    d = _one_json_line(["--steps", "1", "--warmup", "3", "--size", "128", "--count", "8",
                        "--no-himeno", "--no-cpu", "--no-strong"])
    for k in COMMON + ["gpu_launches", "roofline", "clocks"]:
        assert k in d, k
    assert "impl" not in d and d["n_gpus"] == 1 and d["dtype"] == "f32" and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["value"] > 0 and d["gpu_launches"] > 0
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] == 128 ** 3 * 4 == e["d2h_bytes_per_step"]
    r = d["roofline"]
    for k in ["bound", "achieved", "peak", "unit", "frac", "traffic", "kernel"]:
        assert k in r, k
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    # the post-run parity cases against the oracle and the exact-solution checksum
    assert d["parity_ok"] is True, d["parity"]
    assert d["e2e"]["checksum_ok"] is True
    assert len(d["parity"]["bit_exact_cases"]) >= 7 and not any("MISMATCH" in c for c in d["parity"]["bit_exact_cases"])
    assert d["config1_256"]["iter1_loop"]["plan_cache_hits"] > 0

"""The driver's contract with bench.py: one JSON line per arm with the agreed keys.  The reference arm runs here (no GPU,
a tiny sample); the b200 arm runs on the GPU box with a small batch."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"}


def _run(args, timeout):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=timeout,
                       cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [l for l in r.stdout.strip().split("\n") if l.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_prints_the_contract_line(oracle):
    j = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-blocks", "192", "--ref-compress-blocks", "96"], 600)
    assert j["impl"] == "reference" and BASE_KEYS <= set(j)
    assert j["unit"] == "GB/s" and j["value"] > 0 and j["higher_is_better"] is True and j["dtype"] == "u8"
    assert "workload" in j["config"] and j["vs_baseline"] is None
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1 and j["cpu_baseline"]["value"] == j["value"]
    assert j["e2e"] == {"value": j["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    c = j["compress"]  # the other half of BASELINE's metric
    assert c["value"] > 0 and c["cpu_baseline"]["kind"] == "port" and 0.3 < c["config"]["ratio"] < 0.7
    assert j["gpu_launches"] == 0


@pytest.mark.gpu
def test_b200_arm_prints_the_contract_line():
    j = _run(["--gpus", "1", "--steps", "2", "--warmup", "3", "--blocks", "8192", "--e2e-blocks", "2048", "--cpu-blocks", "512",
              "--cpu-seconds", "0.5"], 900)
    assert BASE_KEYS | {"roofline", "clocks", "compress"} <= set(j) and "impl" not in j
    assert j["n_gpus"] == 1 and j["steps"] == 2 and j["warmup"] == 3 and j["dtype"] == "u8" and j["data"] == "synthetic"
    for sec in (j, j["compress"]):
        r = sec["roofline"]
        assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["peak"] > 1000 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3
        assert "k_" in r["kernel"]
        assert sec["cpu_baseline"]["kind"] == "port" and sec["cpu_baseline"]["value"] > 0
        e = sec["e2e"]
        assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
        assert sec["gpu_launches"] > 0 and sec["value"] > 0
    assert j["e2e"]["compressed_in_device_out"]["d2h_bytes_per_step"] == 0
    assert j["gpu_launches"] >= j["steps"] + j["compress"]["gpu_launches"]
    assert set(j["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}

"""The driver's bench contract: ONE JSON line on stdout with the agreed keys, for both arms."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"}


def run_bench(*args):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + list(args), capture_output=True, text=True,
                       cwd=ROOT, timeout=900)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, "stdout must carry exactly one line, got %d" % len(lines)
    return json.loads(lines[0])


def test_reference_arm_line(oracle_built):
    from oracle.ref import reference_available
    if not reference_available("libmultiray_ref.so"):
        pytest.skip("oracle/_ref/libmultiray_ref.so not present")
    d = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "air->ice launch-angle solves/sec" and d["unit"] == "solves/s" and d["higher_is_better"] is True
    assert d["value"] > 1e3 and d["vs_baseline"] is None and "workload" in d["config"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
    assert d["gpu_launches"] == 0


@pytest.mark.gpu
def test_our_arm_line():
    d = run_bench("--steps", "3", "--warmup", "3", "--pairs", "500000", "--skip-extras", "--skip-cpu")
    assert BASE_KEYS | {"clocks", "gpu_launches", "roofline"} <= set(d) and "impl" not in d
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] >= 3 and d["scaling"] == "weak" and d["dtype"] == "f64"
    assert d["value"] > 1e8 and d["gpu_launches"] == 3 and d["vs_baseline"] is None
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and 0 < r["frac"] < 1
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == 16 * 500000 and e["d2h_bytes_per_step"] == 73 * 500000 and 0 < e["value"] < d["value"]
    assert e["matches_device_path"] is True
    c = d["clocks"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(c)

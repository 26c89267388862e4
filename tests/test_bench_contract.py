"""bench.py keeps the driver's contract: one JSON line with the agreed keys.  The reference arm (the oracle port on
the host cores) runs anywhere; the B200 arm needs a GPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e"}


def run_bench(*args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], stdout=subprocess.PIPE,
                         stderr=subprocess.PIPE, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line(monkeypatch):
    """`--impl reference` times the CPU restatement of the path on the host cores with the B200 arm's metric, unit
    and config, and says so (impl, cpu_baseline.kind / cores / sample, e2e without copies)."""
    d = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "sim transitions/s" and d["unit"] == "transitions/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == (os.cpu_count() or 1)
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["sample"]
    assert d["e2e"] == dict(value=d["value"], unit=d["unit"], h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    assert "workload" in d["config"] and "model" not in d["config"]


@pytest.mark.gpu
def test_b200_arm_line():
    d = run_bench("--steps", "3", "--warmup", "3", "--no-cpu-baseline")
    assert BASE_KEYS | {"roofline", "gpu_launches", "clocks", "back_to_back"} <= set(d)
    assert d["n_gpus"] == 1 and d["scaling"] == "weak" and d["data"] == "synthetic" and d["gpu_launches"] > 0
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r)
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12 and 0.2 < r["frac"] < 1.0
    e = d["e2e"]
    # (e2e runs 50 steps without the L2 flush between them: it may come out a little above the flushed, 3-step value)
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] <= 1.15 * d["value"]
    # the per-step windows hide nothing: the same steps back to back, one event pair, within a few percent
    assert abs(d["back_to_back"]["value"] / d["value"] - 1.0) < 0.1
    assert d["clocks"]["sm_mhz"] > 0


@pytest.mark.gpu
@pytest.mark.parametrize("workload", ["c2", "c5", "c3"])
def test_other_baseline_configs_through_bench(workload):
    """`--workload c2 | c5 | c3` put BASELINE configs 2, 5 and 3 behind the same JSON contract (c2 and c5 through the
    reference's Python API, c3 through the C ABI), and the reference arm answers for them too."""
    d = run_bench("--workload", workload, "--steps", "3", "--warmup", "3", "--no-cpu-baseline", "--soak", "0")
    assert BASE_KEYS | {"roofline", "gpu_launches", "clocks"} <= set(d)
    assert d["metric"] == "sim transitions/s" and d["value"] > 1e7 and d["gpu_launches"] > 0
    assert workload in d["config"]["workload"][:3]
    assert d["e2e"]["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    if workload == "c2":
        assert d["rollouts_per_s"] > 1e6 and len(d["last_decision"]) >= 1
    r = run_bench("--impl", "reference", "--workload", workload, "--steps", "1", "--warmup", "0")
    assert r["impl"] == "reference" and r["config"]["workload"] == d["config"]["workload"] and r["unit"] == d["unit"]

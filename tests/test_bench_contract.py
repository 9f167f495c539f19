"""bench.py prints ONE JSON line with the contract's keys (driver contract + tier additions: roofline, cpu_baseline,
e2e, clocks, gpu_launches).  CPU tier: the reference arm on a tiny bounded sample; gpu tier: the CUDA arm on a
200 000-agent population."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e"}


def _run(args, timeout):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, cwd=ROOT, capture_output=True,
                         text=True, timeout=timeout)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.strip().split("\n") if l.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-sample", "20000", "--cpu-workers", "2"], 300)
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "agent_days_per_sec" and d["unit"] == "agent-days/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == 2 and cb["value"] == d["value"] and "sample" in cb
    assert "workload" in d["config"] and "model" not in d["config"]


@pytest.mark.gpu
def test_gpu_arm_line():
    d = _run(["--n", "200000", "--steps", "2", "--warmup", "3", "--no-cpu-baseline", "--c3-n", "300000"], 900)
    assert BASE_KEYS | {"clocks", "gpu_launches", "roofline"} <= set(d) and "impl" not in d
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3 and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["value"] > 0 and d["gpu_launches"] >= 2 * 3
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["peak"] > 0 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert d["e2e"]["value"] > 0 and d["e2e"]["value"] != d["value"] and d["e2e"]["d2h_bytes_per_step"] > 0
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert "workload" in d["config"] and "model" not in d["config"]
    # every BASELINE.json config is a leg of the line, with its own numbers
    cfg = d["config"]
    for leg in ("c1", "c4_like", "c5_like", "c3_single_gpu", "stress_variant", "as_reference_tracing", "dropin_run_model"):
        assert leg in cfg and "error" not in cfg[leg], (leg, cfg.get(leg))
    for leg in ("c1", "c4_like", "c5_like", "c3_single_gpu"):
        assert cfg[leg]["agent_days_per_sec"] > 0 and cfg[leg]["sim_runs_per_hour_e2e"] > 0 and 0 < cfg[leg]["roofline_frac"] < 1.5
    assert cfg["dropin_run_model"]["warm_call_ms"] < cfg["dropin_run_model"]["first_call_ms"]
    k = r["kernels"]
    assert abs(k["day_loop"]["algorithmic_bytes"] + k["seir_materialize_kernel"]["algorithmic_bytes"]
               - r["algorithmic_bytes_per_launch"]) < 1e-6 * r["algorithmic_bytes_per_launch"]

"""bench.py's JSON-line contract: the reference arm (CPU oracle port) here, our arm on a small plate on the GPU box."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"}


def run_bench(*args, timeout=600):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line():
    d = run_bench("--impl", "reference", "--steps", "2", "--warmup", "1", "--ref-nodes", "60", "--ref-iters", "5")
    assert BASE_KEYS <= set(d)
    assert d["impl"] == "reference" and d["metric"] == "CG DOF-iterations/s" and d["unit"] == "DOF-iterations/s"
    assert d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0 and d["higher_is_better"] is True
    assert d["dtype"] == "f64" and d["vs_baseline"] is None and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_stdout_carries_only_the_json_line():
    """whatever a library prints to file descriptor 1 after claim_stdout() lands on stderr"""
    code = ("import bench, os; fd = bench.claim_stdout(); print('banner from a library'); os.write(1, b'raw write\\n'); "
            "bench.emit_line(fd, {'metric': 'x', 'value': 1})")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120, cwd=ROOT)
    assert r.returncode == 0, r.stderr
    assert r.stdout == '{"metric": "x", "value": 1}\n'
    assert "banner from a library" in r.stderr and "raw write" in r.stderr


@pytest.mark.gpu
def test_our_arm_line_on_a_small_plate():
    d = run_bench("--nodes", "200", "--steps", "2", "--warmup", "3", "--iters", "20", "--cpu-iters", "3", "--tts", "on")
    assert BASE_KEYS | {"roofline", "clocks", "metrics", "assembly_roofline"} <= set(d)
    assert "impl" not in d and d["n_gpus"] == 1 and d["scaling"] == "weak" and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["value"] > 0 and d["e2e"]["value"] > 0 and d["e2e"]["value"] < d["value"] * 1.05
    assert d["e2e"]["h2d_bytes_per_step"] == 48 * 200 * 200 and d["e2e"]["d2h_bytes_per_step"] == 48 * 200 * 200
    rf = d["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-12
    assert rf["kernel"] in ("k_spmv_sell", "k_spmv") and rf["bytes_per_launch"] > 0
    assert d["gpu_launches"] == (3 + 3 * 20) * 2
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    tts = d["metrics"]["time_to_solution"]["multilevel"]
    assert tts["converged"] and tts["rel_residual"] <= 1e-8
    assert d["metrics"]["elements_assembled_per_s"] > 0
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}

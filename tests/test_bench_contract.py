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
    d = run_bench("--impl", "reference", "--steps", "2", "--warmup", "1", "--ref-nodes", "60", "--ref-iters", "5", "--tts", "off")
    assert BASE_KEYS <= set(d)
    import bench
    assert d["config"] == bench.workload_config(60, 60, 1, 200)        # the reference arm states OUR arm's config
    assert d["metrics"]["iters_run_per_step"] == 5 and d["metrics"]["elements_assembled_per_s"] > 0
    assert d["impl"] == "reference" and d["metric"] == "CG DOF-iterations/s" and d["unit"] == "DOF-iterations/s"
    assert d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0 and d["higher_is_better"] is True
    assert d["dtype"] == "f64" and d["vs_baseline"] is None and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_ignores_torchruns_thread_cap_and_loads_no_product_code():
    """torchrun exports OMP_NUM_THREADS=1; the reference arm must still use every core it may run on, and must not
    load the product library (its mesh comes from the oracle's own meshGen restatement)"""
    env = dict(os.environ, OMP_NUM_THREADS="1")
    code = ("import sys, json, io, contextlib; sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '0', '--ref-nodes', '40', "
            "'--ref-iters', '3', '--tts', 'off']; import bench; buf = io.StringIO();\n"
            "with contextlib.redirect_stdout(buf): bench.main()\n"
            "d = json.loads(buf.getvalue()); import os; "
            "print(json.dumps({'cores': d['cpu_baseline']['cores'], 'want': len(os.sched_getaffinity(0)), "
            "'product_loaded': any('libfemshell' in l for l in open('/proc/self/maps')), 'fsb_imported': 'fem_shell_b200' in sys.modules}))")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["cores"] == d["want"] and not d["product_loaded"] and not d["fsb_imported"], d


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
    assert rf["kernel"] in ("k_spmv_sell", "k_spmv") and rf["bytes_per_launch"] > 0 and "traffic_source" in rf
    import bench
    assert d["config"] == bench.workload_config(200, 200, 1, 20)
    first = d["metrics"]["time_to_first_solution"]
    assert first["converged"] and first["seconds"] > 0
    small = d["metrics"]["time_to_solution_bounded"]
    assert small["converged"] and small["iterations"] > 100
    assert d["cpu_baseline"]["time_to_solution_bounded"]["iterations"] > 100
    assert d["gpu_launches"] == (3 + 3 * 20) * 2
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    tts = d["metrics"]["time_to_solution"]["multilevel"]
    assert tts["converged"] and tts["rel_residual"] <= 1e-8
    assert d["metrics"]["elements_assembled_per_s"] > 0
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}

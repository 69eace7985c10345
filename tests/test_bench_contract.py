"""bench.py's one-line JSON contract: the reference arm and the no-GPU behaviour on the CPU, our arm on a B200."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT, have_gpu

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config"}


def run_bench(*args):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, cwd=ROOT)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    return r, (json.loads(lines[-1]) if lines else None)


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "sipnet_ref")), reason="oracle/_ref not built")
def test_reference_arm_line():
    r, line = run_bench("--impl", "reference", "--steps", "1", "--warmup", "1", "--members", "64", "--years", "1")
    assert r.returncode == 0 and line is not None, r.stderr[-2000:]
    assert BASE_KEYS <= set(line) and line["impl"] == "reference"
    assert line["metric"] == "ensemble member-timesteps/sec" and line["unit"] == "member-timesteps/s"
    assert line["higher_is_better"] is True and line["vs_baseline"] is None and line["dtype"] == "f64"
    assert "workload" in line["config"] and "model" not in line["config"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["value"] > 1e4


@pytest.mark.skipif(have_gpu(), reason="needs a machine WITHOUT a GPU")
def test_our_arm_refuses_to_run_without_a_gpu():
    r, line = run_bench("--steps", "1", "--warmup", "1", "--members", "64", "--years", "1")
    assert r.returncode != 0 and line is None
    assert "no CPU fallback" in (r.stderr + r.stdout)


@pytest.mark.gpu
def test_our_arm_line():
    r, line = run_bench("--steps", "2", "--warmup", "3", "--members", "2048", "--years", "1", "--no-cpu-baseline")
    assert r.returncode == 0 and line is not None, r.stderr[-2000:]
    assert BASE_KEYS <= set(line) and "impl" not in line
    assert line["n_gpus"] == 1 and line["steps"] == 2 and line["warmup"] == 3 and line["scaling"] == "weak"
    assert line["comm_nranks"] == 1 and "C4" in line["config"]["workload"]
    roof = line["roofline"]
    assert roof["bound"] in ("fp64", "hbm") and roof["unit"] in ("TFLOP/s", "GB/s")
    assert abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-12 and 0 < roof["frac"] < 1
    assert set(line["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert line["gpu_launches"] >= 4 * line["steps"]              # init_state, step kernel, replay, summary kernels per pass
    e2e = line["e2e"]
    T = line["config"]["model_steps"]
    assert e2e["h2d_bytes_per_step"] == 80 * 2048 * 8 and e2e["d2h_bytes_per_step"] == (2 + 2 + 6) * T * 8
    assert 0 < e2e["value"] < line["value"]                       # the copies are inside the timed region
    assert line["oracle_spot_check"]["bit_identical_to_oracle"] is True
    assert line["c5"]["comm_nranks"] == 1 and line["c5"]["value"] > 0
    assert line["c2"]["e2e"]["d2h_bytes_per_step"] == 32 * T * 4096 * 8
    assert line["c2"]["d2h_probe"]["gb_s_all_ranks"] >= line["c2"]["e2e"]["d2h_gb_s_all_ranks"] > 0   # the ceiling beside it


def test_both_arms_share_one_config_object():
    """`config` must be the same dict in both arms (the driver compares them)."""
    import argparse
    sys.path.insert(0, ROOT)
    import bench
    args = argparse.Namespace(members=131072, years=10)
    a = bench.bench_config(args, 7306)
    assert a == bench.bench_config(args, 7306) and set(a) >= {"workload", "members_per_gpu", "model_steps"}
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count('"config": bench_config(args, T)') == 2      # our arm and the reference arm


def test_host_writer_section_runs_without_a_gpu():
    """bench.py's informational `host_writer` section is host C only (the drop-in driver's main-output writer)."""
    sys.path.insert(0, ROOT)
    import bench
    r = bench.host_writer_rate(300, members=19)
    assert r["value"] > 0 and r["text_gb_s"] > 0 and r["threads"] >= 1 and "19 members x 300 steps" in r["sample"]

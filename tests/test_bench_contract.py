"""bench.py output contract: exactly one JSON line on stdout with the keys the driver reads, for the reference arm
(CPU, runs anywhere oracle/_ref is built) and for the GPU arm."""
import json
import os
import subprocess
import sys

import pytest

import oracle_ref

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e"}


def run_bench(*flags):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *flags], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines                      # nothing but the JSON line on stdout
    return json.loads(lines[0])


@pytest.mark.skipif(not oracle_ref.available(), reason="oracle/_ref not built")
def test_reference_arm_line():
    d = run_bench("--impl", "reference", "--steps", "1", "--warmup", "1", "--tran-budget", "1")
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    t = d["tran_c3"]                                    # .TRAN leg: BASELINE config 3 on the host cores
    assert t["kind"] == "reference" and t["cores"] >= 1 and t["wall_s"] > 0 and t["rings_run"] >= t["cores"]
    assert 20 < t["accepted_steps_per_ring"] < 200 and t["newton_iters_per_ring"] > t["accepted_steps_per_ring"]
    assert d["config"]["instances_per_gpu"] == 100000   # the full array, not a sample
    assert d["metric"] == "bsim4_device_load_stamp_evals_per_sec" and d["unit"] == "evals/s" and d["dtype"] == "f64"
    assert d["value"] > 1e5 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


@pytest.mark.gpu
def test_gpu_arm_line():
    d = run_bench("--steps", "4", "--warmup", "3", "--no-cpu-baseline", "--no-tran")
    assert BASE_KEYS | {"roofline", "gpu_launches", "clocks"} <= set(d)
    assert d["metric"] == "bsim4_device_load_stamp_evals_per_sec" and d["n_gpus"] == 1 and d["steps"] == 4
    assert d["value"] > 1e8 and d["gpu_launches"] >= 2
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    e = d["e2e"]
    assert 0 < e["value"] < d["value"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert "workload" in d["config"] and "l2" in d["config"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])


@pytest.mark.skipif(not oracle_ref.available(), reason="oracle/_ref not built")
def test_reference_arm_under_torchrun_prints_once():
    """launched like the driver does for N > 1: rank 0 alone runs and prints the line, the other ranks exit 0"""
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "1", "--no-tran"], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 1e5

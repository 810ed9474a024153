"""bench.py's reference arm runs on the CPU (the oracle port timed on the host cores) and must print ONE JSON line with
the contract's keys; the GPU arm's line is checked on the B200 (`-m gpu`)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"}


def _one_line(args, timeout):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=timeout,
                       cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, p.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_prints_the_contract_line():
    line = _one_line(["--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-seconds", "2"], 600)
    assert BASE_KEYS <= set(line), BASE_KEYS - set(line)
    assert line["impl"] == "reference" and line["unit"] == "photons/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["cpu_baseline"]["as_shipped"]["value"] <= line["cpu_baseline"]["value"]
    assert "workload" in line["config"] and "model" not in line["config"]
    # both arms emit the SAME config object (the driver compares them)
    sys.path.insert(0, ROOT)
    import bench

    class A:
        views = False; workload = "c3"
    assert line["config"] == bench.config_of(A)


@pytest.mark.gpu
def test_gpu_arm_prints_the_contract_line():
    line = _one_line(["--steps", "1", "--warmup", "3", "--photons", "4000000", "--no-cpu-baseline"], 900)
    assert (BASE_KEYS - {"cpu_baseline"}) | {"clocks", "gpu_launches", "roofline"} <= set(line)
    assert line["n_gpus"] == 1 and line["gpu_launches"] == 1 and line["value"] > 1e8
    r = line["roofline"]
    assert r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["unit"] == "GB/s"
    g = r["l2_gather"]                                    # the binding ceiling, measured in the same run
    assert g["peak"] > 1e11 and 0.2 < g["frac"] < 1.5 and abs(g["frac"] - g["achieved"] / g["peak"]) < 1e-9
    assert g["measured_ceilings_gathers_per_s"]["hbm_1GB"] < g["measured_ceilings_gathers_per_s"]["l2_16MB"]
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0
    assert line["e2e"]["value"] != line["value"]
    assert set(line["config"]) == {"workload", "l2"}
    assert line["details"]["bad_photons"] == 0
    f = line["details"]["fluxes"]
    assert abs(f["meanFluxUp"] + f["meanFluxDown"] + f["meanFluxAbsorbed"] - 1.0) < 2e-3      # albedo 0: energy closure

"""bench.py's reference arm on the CPU (it needs no GPU): the JSON line the driver parses, on a small workload."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "covid_c",
                          "--steps", "2", "--warmup", "1", "--ref-envs-per-core", "2"], capture_output=True, text=True, timeout=600,
                         cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "env-steps/s" and line["unit"] == "env-steps/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["steps"] == 2 and line["n_gpus"] == 1
    assert "COVID_C" in line["config"]["workload"] and line["config"]["scenario"] == "covid_c"
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "oracle/epi_oracle.c" in cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0
    # no GPU library may be mapped by this arm (the driver records the .so files the process loaded)
    assert "libepi_b200" not in out.stderr

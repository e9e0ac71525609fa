"""bench.py's contract, as far as it can be checked without a GPU: the reference arm (`--impl reference`) runs the
reference's own CPU code and prints exactly ONE JSON line on stdout (the reference's std::cout chatter goes to
stderr), with the keys and the config the driver compares with the GPU arm's; the GPU arm fails loudly without a
device instead of falling back to anything."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                          "--ref-sample-particles", "20000", "--no-mpi-baseline"], capture_output=True, text=True, timeout=600,
                         cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, res.stdout[:2000]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "attempted_mc_moves_per_s" and d["unit"] == "moves/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["gpu_launches"] == 0
    assert d["e2e"] == {"value": d["value"], "unit": "moves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] == 1 and cb["value"] == d["value"] and cb["sample"]
    # the same workload description as the GPU arm's line
    import bench

    assert d["config"] == bench.config_dict(1, bench.default_per_gpu(1), "LennardJones", 5)
    assert d["config"]["workload"] == "gcmc_lj_1000k_per_gpu"


def test_gpu_arm_refuses_to_run_without_a_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"], capture_output=True,
                         text=True, timeout=600, cwd=ROOT)
    assert res.returncode != 0
    assert not any(l.strip().startswith("{") for l in res.stdout.splitlines())  # no number without a GPU

"""bench.py's JSON contract on the CPU: the reference arm (`--impl reference`, the CPU oracle on the host cores) is the
one leg that runs without a GPU.  Checks the keys the driver reads, that ranks other than 0 stay silent under a
multi-rank launch, and that the GPU arm refuses to start without a CUDA device instead of falling back."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None, timeout=300):
    e = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        e.pop(k, None)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, cwd=ROOT,
                          env=e, timeout=timeout)


@pytest.mark.timeout(600)
def test_reference_arm_line_has_the_contract_keys():
    res = _run(["--impl", "reference", "--batch", "1", "--size", "64", "--steps", "1", "--warmup", "1", "--gpus", "1"])
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["metric"] == "CycleGAN train img/s" and line["unit"] == "img/s"
    assert line["higher_is_better"] is True and line["vs_baseline"] is None and line["data"] == "synthetic"
    assert line["value"] > 0 and abs(line["value"] - 1e3 / line["ms_per_step"]) < 1e-6 * line["value"] + 1e-9
    assert "workload" in line["config"] and "model" not in line["config"]
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_are_silent():
    res = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_gpu_arm_does_not_fall_back_to_the_cpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a box without a CUDA device")
    res = _run(["--steps", "1", "--warmup", "0", "--no-cpu-baseline", "--no-roofline"])
    assert res.returncode != 0 and res.stdout.strip() == ""   # no line: there is no CPU compute path to measure

"""bench.py contract, the part that runs without a GPU: the reference arm (CPU restatement on the host cores) prints ONE
JSON line with the keys the driver reads; under torchrun only rank 0 prints; the GPU arm refuses to run without a device
instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

from conftest import ROOT

BENCH = os.path.join(ROOT, "bench.py")


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, BENCH] + args, capture_output=True, text=True, env=e, timeout=600)


def test_reference_arm_prints_one_contract_line():
    r = _run(["--impl", "reference", "--steps", "2", "--warmup", "1", "--scale", "0.02"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "ms/LM-iter" and d["higher_is_better"] is False
    assert d["metric"] == json.load(open(os.path.join(ROOT, "BASELINE.json")))["metric"]
    assert d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0 and d["ms_per_step"] == d["value"]
    assert d["config"]["workload"].startswith("synthetic ladybug-1723")
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": "ms/LM-iter", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert len(d["cg_iterations"]) == 2
    one = d["cpu_baseline_1thread"]  # T = all cores and T = 1
    assert one["cores"] == 1 and one["value"] > 0 and one["unit"] == "ms/LM-iter"


def test_reference_arm_other_ranks_stay_silent():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "1", "--scale", "0.02", "--gpus", "2"], env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_has_no_cpu_fallback():
    r = _run(["--steps", "1", "--warmup", "1", "--scale", "0.02"], env={"CUDA_VISIBLE_DEVICES": ""})
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)

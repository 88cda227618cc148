"""bench.py's reference arm (`--impl reference`) runs on the host cores alone: it must work without a
GPU and print ONE JSON line carrying the contract's keys.  (The B200 arm needs a GPU and refuses to
run without one — no CPU fallback.)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, env=e, cwd=ROOT,
                          timeout=600)


def test_reference_arm_json_line(built):
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--width", "96", "--height", "54")
    assert r.returncode == 0, r.stderr
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "ray_steps_per_sec" and d["unit"] == "ray-steps/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and "workload" in d["config"]


def test_reference_arm_other_ranks_exit_quietly(built):
    r = _run("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_b200_arm_refuses_to_run_without_a_gpu(built):
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is visible")
    r = _run("--steps", "1", "--warmup", "0")
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


def test_bench_kernel_labels_have_an_ncu_record():
    """bench.py looks its ncu figures (roofline.traffic, fp64_pipe) up by the kernel label: a label without a record in
    profiles/latest_traffic.json would silently drop them from the line."""
    import json, os, re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "bench.py")).read()
    m = re.search(r'KERNEL = \{"f64_fast": "([^"]+)", "f64": "([^"]+)"\}', src)
    assert m, "bench.py: KERNEL labels not found"
    prof = json.load(open(os.path.join(root, "profiles", "latest_traffic.json")))
    for label in m.groups():
        assert label in prof, label
        assert prof[label]["dram_bytes_per_launch"] > 0 and os.path.exists(os.path.join(root, prof[label]["source"].split(" ")[0]))

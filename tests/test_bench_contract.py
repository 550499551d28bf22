"""bench.py contract pieces that can be checked without a GPU: the reference arm (the reference algorithm on the host
cores, oracle/fgd_mirror.c) prints the contract's JSON line, and the B200 arm refuses to run without a CUDA device
instead of falling back to anything."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=240):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, env=env, capture_output=True,
                          text=True, timeout=timeout)


def test_reference_arm_prints_the_contract_line():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-seconds", "2", "--batch", "512")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "optimized trajectories/sec" and d["unit"] == "trajectories/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f32" and d["data"] == "synthetic"
    assert d["value"] > 0 and d["steps"] == 1 and d["n_gpus"] == 1
    assert d["config"]["workload"].startswith("c5") and "model" not in d["config"]      # the north-star configuration by default
    assert d["scaling"] == "strong" and d["config"]["trajectories_total"] == 512
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "oracle/fgd_mirror.c" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_b200_arm_has_no_cpu_fallback():
    r = _run("--steps", "1", "--warmup", "0", timeout=120)
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stderr + r.stdout)

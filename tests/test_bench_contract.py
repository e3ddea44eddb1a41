"""CPU: the benchmark contract that can be checked without a GPU -- the reference arm prints one
JSON line with the keys the driver reads, non-zero ranks of a torchrun launch print nothing, and
the GPU arm refuses to run without a device instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, BENCH] + args, capture_output=True, text=True, env=e, timeout=600)


def test_reference_arm_line():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "3", "--npts", "10000"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "mpc_solves_per_sec" and d["unit"] == "solves/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f64"
    assert d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 3
    assert "workload" in d["config"] and "model" not in d["config"]
    # both arms print the same config dict, key for key (the driver compares them)
    sys.path.insert(0, ROOT)
    import argparse
    import bench
    ns = argparse.Namespace(batch=1024, npts=10000, warm="ref", tol=1e-8, max_iter=50)
    assert d["config"] == bench.shared_config(ns, 1)
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"]
    assert e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_reference_arm_other_ranks_are_silent():
    r = _run(["--impl", "reference", "--steps", "1", "--npts", "10000", "--gpus", "2"],
             env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    r = _run(["--steps", "1"])
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stderr + r.stdout)


def test_cpu_c0_mode_runs_without_a_gpu():
    """BASELINE config C0 (one instance, K = 8, 10k points, one thread) is a CPU measurement: the
    mode must work on a box without a GPU and report per-stage p50 / p90."""
    r = _run(["--mode", "cpu_c0"])
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([l for l in r.stdout.strip().splitlines() if l.startswith("{")][-1])
    assert d["mode"] == "cpu_c0" and d["unit"] == "ms" and d["converged_frac"] == 1.0
    for k in ("tree_build", "knn_20x8", "nlp_solve", "total"):
        assert 0 < d[k]["p50"] <= d[k]["p90"]

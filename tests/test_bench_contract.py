"""bench.py contract checks that need no GPU: the reference arm (`--impl reference`) prints exactly one JSON line with the keys
the driver reads, on a tiny workload; under a multi-rank launch only rank 0 prints."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "12", "--steps", "1",
                        "--warmup", "0"], capture_output=True, env=env, cwd=ROOT, timeout=600)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    return [ln for ln in r.stdout.decode().splitlines() if ln.strip().startswith("{")]


def test_reference_arm_json_line():
    lines = _run()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "MDoF*iter/s" and d["higher_is_better"] is True
    for k in ("metric", "value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config",
              "cpu_baseline", "e2e", "gpu_launches"):
        assert k in d, k
    assert d["value"] > 0 and d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["single_thread_value"] > 0 and cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    assert d["gpu_launches"] == 0


def test_reference_arm_other_ranks_print_nothing():
    assert _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []


def test_stdout_guard_keeps_stdout_to_one_line():
    """bench.py's StdoutGuard: noise written to fd 1 by native code (NCCL's version banner) must not reach stdout"""
    code = (
        "import os, sys\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "import bench\n"
        "g = bench.StdoutGuard()\n"
        "os.write(1, b'NCCL version 2.28.9+cuda12.9\\n')\n"
        "print('python noise')\n"
        "g.emit('{\"ok\": 1}')\n"
        "g.close()\n"
    )
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, cwd=ROOT, timeout=120)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    assert r.stdout.decode() == '{"ok": 1}\n'
    assert b"NCCL version" in r.stderr and b"python noise" in r.stderr


@pytest.mark.gpu
def test_ours_arm_json_line_on_a_small_workload():
    """the product arm on one GPU with the S1 workload (seconds): ONE JSON line on stdout with every key of the contract, the
    roofline / e2e / clocks objects filled in and a non-zero count of our own kernel launches"""
    import subprocess
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "S1", "--steps", "2", "--warmup", "3", "--no-cpu", "--no-extras"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
              "roofline", "clocks", "gpu_launches", "e2e", "true_rel_residual"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["dtype"] == "f64" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["gpu_launches"] > 0 and d["value"] > 0
    assert d["pcg_iters"] == 861 and d["true_rel_residual"] < 1e-8
    rf = d["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9 and 0.2 < rf["frac"] < 1.2
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] <= d["value"] * 1.05
    assert "sm_mhz" in d["clocks"] and "reasons" in d["clocks"]

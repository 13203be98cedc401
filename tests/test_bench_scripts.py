"""benchmark/bench_configs.py dry run on CPU/gloo (``--tiny``): the script that reports the BASELINE.json configs and
their roofline fractions keeps working without a GPU in the loop."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bench_configs_tiny_world2():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29691", os.path.join(ROOT, "benchmark", "bench_configs.py"), "--tiny"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:]
    rows = [json.loads(line) for line in r.stdout.splitlines() if line.startswith("{")]
    assert [x["config"] for x in rows] == ["zigzag", "varlen", "llama3", "stripe8"]
    for x in rows:
        assert x["n_gpus"] == 2 and x["ms_per_step"] > 0
        roof = x["roofline"]
        assert roof["compute_ms"] > 0 and roof["nvlink_ms"] > 0 and roof["bound"] in ("compute", "nvlink")


def test_bench_reference_arm_reports_unavailable_without_a_gpu():
    """The driver contract: `bench.py --impl reference` prints one JSON line and exits 0 when it cannot run."""
    import torch

    if torch.cuda.is_available():
        import pytest

        pytest.skip("CPU-only check")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference"], stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-1000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and "unavailable" in line
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                       text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0 and "CUDA" in (r.stderr + r.stdout)

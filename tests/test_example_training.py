"""examples/train_hf_llama_cp.py end to end on CPU/gloo: the same packed batches give the same loss trajectory
(forward, backward, gradient all-reduce, AdamW) whether the sequence lives on one rank or is split over two."""
import os
import re
import subprocess
import sys

import pytest

pytest.importorskip("transformers")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _losses(world, port, layout="llama3"):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "examples", "train_hf_llama_cp.py"), "--tokens", "96", "--steps", "3",
           "--layout", layout]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:]
    return [float(x) for x in re.findall(r"loss ([0-9.]+)", r.stdout)]


def test_cp_training_matches_single_rank():
    one = _losses(1, 29681)
    two = _losses(2, 29682)
    zig = _losses(2, 29683, "zigzag")
    assert len(one) == 3 and len(two) == 3 and len(zig) == 3
    for a, b, c in zip(one, two, zig):
        assert abs(a - b) < 2e-3 and abs(a - c) < 2e-3, (one, two, zig)

"""NVTX ranges around the context-parallel ops (``RFA_B200_NVTX=1``), for nsys / ncu ``--nvtx`` filtering.

The reference's only tracing is ``torch.profiler`` inside its benchmark scripts
(/root/reference/benchmark/benchmark_kvpacked_func.py:55-80); here ``benchmark/*.py --profile`` does the same, the
kernels can record per-CTA ``clock64`` timelines (``RFA_TRACE``) and every forward / backward op can be
bracketed by a named range: ``rfa.<scheme>.fwd`` / ``rfa.<scheme>.bwd``.
"""
from __future__ import annotations

import contextlib
import os

import torch


def enabled() -> bool:
    return os.environ.get("RFA_B200_NVTX", "0") == "1"


@contextlib.contextmanager
def nvtx(name: str, like: torch.Tensor):
    """Push / pop an NVTX range when enabled and ``like`` lives on a CUDA device; free otherwise."""
    on = enabled() and like.is_cuda
    if on:
        torch.cuda.nvtx.range_push(name)
    try:
        yield
    finally:
        if on:
            torch.cuda.nvtx.range_pop()

#!/bin/bash
# Last seconds of the round-2 GPU budget: smoke() (packed-gradient path) and the deterministic launch groups.
mkdir -p gpurun_out
timeout 80 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/last_smoke.log 2>&1; echo "smoke exit $?"; tail -n 2 gpurun_out/last_smoke.log
timeout 60 python -m pytest tests/test_zz_gpu_deterministic.py -m gpu -x -q > gpurun_out/last_det.log 2>&1; echo "det exit $?"; tail -n 5 gpurun_out/last_det.log

#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 300 -x > gpurun_out/pytest.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest.log
echo "== first look"; timeout 600 python benchmark/first_look.py > gpurun_out/first_look.log 2>&1; echo "first_look exit $?"; tail -4 gpurun_out/first_look.log

#!/bin/bash
# Round-2 single-GPU kernel iteration: correctness of the default kernels, speed, optionally a clock64 trace of both
# kernels from an RFA_TRACE build made on the box (nvcc is in the image; the default build is restored afterwards).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 600 > gpurun_out/pytest_kernels.log 2>&1; echo "kernel tests exit $?"; tail -6 gpurun_out/pytest_kernels.log | cut -c1-200
RFA_FIRST_LOOK_SKIP_FA2=${SKIP_FA2:-1} RFA_FIRST_LOOK_OUT=gpurun_out/first_look_new.json timeout 600 python benchmark/first_look.py > gpurun_out/first_look_new.log 2>&1; echo "first_look exit $?"; cat gpurun_out/first_look_new.log | cut -c1-400
if [ "${TRACE:-0}" = "1" ]; then
  cp ring_flash_attn_b200/_C*.so /tmp/_C_default.so
  RFA_TRACE=1 python -c "from ring_flash_attn_b200 import build_ext; build_ext.build(force=True)" > gpurun_out/trace_build.log 2>&1; echo "trace build exit $?"
  timeout 300 python benchmark/trace_fwd.py > gpurun_out/trace_fwd.log 2>&1; echo "trace fwd exit $?"; tail -12 gpurun_out/trace_fwd.log
  timeout 300 python benchmark/trace_bwd.py > gpurun_out/trace_bwd.log 2>&1; echo "trace bwd exit $?"; tail -12 gpurun_out/trace_bwd.log
  cp /tmp/_C_default.so ring_flash_attn_b200/$(basename ring_flash_attn_b200/_C*.so)
fi

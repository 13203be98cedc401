#!/bin/bash
# Forward softmax tuning sweep (1 GPU): polynomial exp2 share and turn-taking, after the PV->QK pipeline change.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_kernels.log 2>&1; echo "kernel tests exit $?"; tail -4 gpurun_out/pytest_kernels.log
for cfg in "0 0" "1 0" "2 0" "0 1" "1 1" "2 1"; do
  set -- $cfg
  RFA_B200_POLY_EXP=$1 RFA_B200_FWD_FLAGS=$2 RFA_FIRST_LOOK_SKIP_FA2=1 RFA_FIRST_LOOK_OUT=gpurun_out/first_look_poly$1_flags$2.json timeout 300 python benchmark/first_look.py > gpurun_out/first_look_poly$1_flags$2.log 2>&1
  echo "poly=$1 flags=$2: $(grep -o '"ours_fwd_tflops": [0-9.]*' gpurun_out/first_look_poly$1_flags$2.log | tr '\n' ' ')"
done
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2_ours_1.log 2>&1; grep '"metric"' gpurun_out/bench_r2_ours_1.log | cut -c1-1500

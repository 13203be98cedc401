#!/bin/bash
# Round-2 single-GPU validation of the kernel variants written after the last round-1 hardware session
# (run under plain `gpurun`, 1 GPU): sliding window, fp8 (descriptor probe first), 64-key-step forward, direct dQ.
mkdir -p gpurun_out
echo "== experimental kernel variants written after the last round-1 hardware session (sliding window)"
RFA_B200_TEST_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 600 -k "sliding_window_kernels or single_token" > gpurun_out/pytest_window.log 2>&1; echo "exit $?"; tail -5 gpurun_out/pytest_window.log
echo "== fp8: kind::f8f6f4 descriptor probe (sweeps alternatives on a mismatch), then the fp8 forward kernel"
timeout 300 python benchmark/probe_fp8.py > gpurun_out/probe_fp8.log 2>&1; echo "exit $?"; tail -4 gpurun_out/probe_fp8.log
RFA_B200_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 500 -k "fp8" > gpurun_out/pytest_fp8.log 2>&1; echo "exit $?"; tail -5 gpurun_out/pytest_fp8.log
echo "== forward variant with 64-key softmax steps (RFA_B200_FWD_H64=1): correctness, then speed vs the default"
RFA_B200_FWD_H64=1 timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 600 -k "fwd_block or world1 or bwd_block" > gpurun_out/pytest_h64.log 2>&1; echo "exit $?"; tail -5 gpurun_out/pytest_h64.log
for h in 0 1; do
  RFA_B200_FWD_H64=$h RFA_FIRST_LOOK_SKIP_FA2=1 RFA_FIRST_LOOK_OUT=gpurun_out/first_look_h64_$h.json timeout 600 python benchmark/first_look.py > gpurun_out/first_look_h64_$h.log 2>&1; echo "h64=$h exit $?"; grep -i "fwd" gpurun_out/first_look_h64_$h.log | head -8
done
RFA_B200_FWD_H64=1 RFA_B200_FWD_FLAGS=1 RFA_FIRST_LOOK_SKIP_FA2=1 RFA_FIRST_LOOK_OUT=gpurun_out/first_look_h64_noturn.json timeout 600 python benchmark/first_look.py > gpurun_out/first_look_h64_noturn.log 2>&1; echo "h64 without turn-taking exit $?"; grep -i "fwd" gpurun_out/first_look_h64_noturn.log | head -4
for poly in 1 2; do
  RFA_B200_FWD_H64=1 RFA_B200_POLY_EXP=$poly RFA_FIRST_LOOK_SKIP_FA2=1 RFA_FIRST_LOOK_OUT=gpurun_out/first_look_h64_poly$poly.json timeout 600 python benchmark/first_look.py > gpurun_out/first_look_h64_poly$poly.log 2>&1; echo "h64 poly=$poly exit $?"; grep -i "fwd" gpurun_out/first_look_h64_poly$poly.log | head -4
done
echo "== backward variants (RFA_B200_BWD_V2 bit 0: dQ^T via coalesced red.global from registers, bit 1: dS^T as TMEM operand of dK)"
for v in 1 2 3; do
  RFA_B200_BWD_V2=$v timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 600 -k "bwd_block or world1" > gpurun_out/pytest_bwdv2_$v.log 2>&1; echo "v2=$v tests exit $?"; tail -3 gpurun_out/pytest_bwdv2_$v.log
  RFA_B200_BWD_V2=$v RFA_FIRST_LOOK_SKIP_FA2=1 RFA_FIRST_LOOK_OUT=gpurun_out/first_look_bwdv2_$v.json timeout 600 python benchmark/first_look.py > gpurun_out/first_look_bwdv2_$v.log 2>&1; echo "v2=$v bench exit $?"; grep -i "bwd" gpurun_out/first_look_bwdv2_$v.log | head -4
done

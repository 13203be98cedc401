#!/bin/bash
# Round-2 first measurement (N = all visible GPUs): where does the MHA headline lose time?
#   - validates the experimental bulk dK/dV epilogue (RFA_B200_DKV_BULK=1) against the oracle,
#   - times fwd and fwd+bwd for the MHA headline shard and the GQA README shard with the epilogue off/on
#     and a few push-CTA counts.
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== correctness with bulk dK/dV epilogue"
RFA_B200_DKV_BULK=1 timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 500 -k "2gpu and True" > gpurun_out/pytest_dkv_bulk.log 2>&1; echo "exit $?"; tail -3 gpurun_out/pytest_dkv_bulk.log
for bulk in 0 1; do
  echo "== breakdown DKV_BULK=$bulk"
  RFA_B200_DKV_BULK=$bulk SWEEP=16,24,32 timeout 600 $TR --master-port 29542 benchmark/multi_breakdown.py > gpurun_out/breakdown_bulk$bulk.log 2>&1
  grep -E "^t[0-9]" gpurun_out/breakdown_bulk$bulk.log | cut -c1-120
done
echo "== BASELINE.json configs 2-5 (ours, then the reference), roofline fractions"
for impl in ours reference; do
  timeout 900 $TR --master-port 29543 benchmark/bench_configs.py --impl $impl > gpurun_out/bench_configs_${impl}_$N.jsonl 2> gpurun_out/bench_configs_${impl}_$N.err
  grep '^{' gpurun_out/bench_configs_${impl}_$N.jsonl | cut -c1-260
done
echo "== experimental kernel variants written after the last round-1 hardware session (sliding window)"
RFA_B200_TEST_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_multi.py -m gpu -q --timeout 600 -k "sliding_window_kernels" > gpurun_out/pytest_window.log 2>&1; echo "exit $?"; tail -5 gpurun_out/pytest_window.log
echo "== fp8: kind::f8f6f4 descriptor probe (sweeps alternatives on a mismatch), then the fp8 forward kernel"
timeout 300 python benchmark/probe_fp8.py > gpurun_out/probe_fp8.log 2>&1; echo "exit $?"; tail -4 gpurun_out/probe_fp8.log
RFA_B200_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 500 -k "fp8" > gpurun_out/pytest_fp8.log 2>&1; echo "exit $?"; tail -5 gpurun_out/pytest_fp8.log
echo "== forward variant with 64-key softmax steps (RFA_B200_FWD_H64=1): correctness, then speed vs the default"
RFA_B200_FWD_H64=1 timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 600 -k "fwd_block or world1 or bwd_block" > gpurun_out/pytest_h64.log 2>&1; echo "exit $?"; tail -5 gpurun_out/pytest_h64.log
for h in 0 1; do
  RFA_B200_FWD_H64=$h RFA_FIRST_LOOK_SKIP_FA2=1 RFA_FIRST_LOOK_OUT=gpurun_out/first_look_h64_$h.json timeout 600 python benchmark/first_look.py > gpurun_out/first_look_h64_$h.log 2>&1; echo "h64=$h exit $?"; grep -i "fwd" gpurun_out/first_look_h64_$h.log | head -8
done
echo "== backward variant: dQ^T added with coalesced red.global straight from registers (RFA_B200_DQ_DIRECT=1)"
RFA_B200_DQ_DIRECT=1 timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 600 -k "bwd_block or world1" > gpurun_out/pytest_dqdirect.log 2>&1; echo "exit $?"; tail -5 gpurun_out/pytest_dqdirect.log
RFA_B200_DQ_DIRECT=1 RFA_FIRST_LOOK_SKIP_FA2=1 RFA_FIRST_LOOK_OUT=gpurun_out/first_look_dqdirect.json timeout 600 python benchmark/first_look.py > gpurun_out/first_look_dqdirect.log 2>&1; echo "exit $?"; grep -i "bwd" gpurun_out/first_look_dqdirect.log | head -8

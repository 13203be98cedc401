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
echo "== sliding-window kernel variants inside the fused 2-GPU launch"
RFA_B200_TEST_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 600 -k "sliding_window_kernels" > gpurun_out/pytest_window_multi.log 2>&1; echo "exit $?"; tail -5 gpurun_out/pytest_window_multi.log
echo "== stripe fp8 config, forward only: dequantise path vs fp8 forward kernel (per-head descales)"
for lvl in 0 1 2; do
  RFA_B200_FP8_KERNEL=$lvl timeout 600 $TR --master-port 29544 benchmark/bench_configs.py --only stripe8 --forward-only --fp8-per-head > gpurun_out/bench_stripe8_fp8kernel$lvl.jsonl 2> gpurun_out/bench_stripe8_fp8kernel$lvl.err
  grep '^{' gpurun_out/bench_stripe8_fp8kernel$lvl.jsonl | cut -c1-220
done
if [ "$N" -ge 4 ]; then
  echo "== two fused context-parallel subgroups side by side (first hardware run)"
  RFA_B200_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 500 -k "fused_subgroups" > gpurun_out/pytest_subgroups.log 2>&1; echo "exit $?"; tail -3 gpurun_out/pytest_subgroups.log
fi

#!/bin/bash
# Round-2 multi-GPU validation (N = all visible GPUs): fused tests (model-dtype dK/dV inbox, device-side needs
# exchange for llama3, window kernels, compiled schemes), concurrency stress, per-phase breakdown, headline bench
# with the sampled oracle check; optionally the reference arm and the BASELINE.json configs for both arms.
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
export RFA_B200_PEER_TIMEOUT_S=${RFA_B200_PEER_TIMEOUT_S:-30}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "${TESTS:-1}" = "1" ]; then
  timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 300 ${PYTEST_K:+-k "$PYTEST_K"} > gpurun_out/pytest_multi_$N.log 2>&1; echo "multi tests exit $?"; tail -8 gpurun_out/pytest_multi_$N.log | cut -c1-200
fi
STRESS_ITERS=${STRESS_ITERS:-1000} timeout 300 $TR --master-port 29547 benchmark/stress_overlap.py > gpurun_out/stress_overlap_$N.log 2>&1; echo "stress exit $?"; grep -E '^\{|differs|rfa:' gpurun_out/stress_overlap_$N.log | head -5
SWEEP=24 timeout 300 $TR --master-port 29542 benchmark/multi_breakdown.py > gpurun_out/breakdown_r2_$N.log 2>&1; grep -E "^t[0-9]" gpurun_out/breakdown_r2_$N.log | cut -c1-140
export RFA_B200_PEER_TIMEOUT_S=600
timeout 600 $TR --master-port 29545 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_r2_ours_$N.log 2>&1; grep '"metric"' gpurun_out/bench_r2_ours_$N.log | cut -c1-1800
if [ "${REF:-0}" = "1" ]; then
  timeout 600 $TR --master-port 29546 bench.py --impl reference --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_r2_reference_$N.log 2>&1; grep '"metric"' gpurun_out/bench_r2_reference_$N.log | cut -c1-700
fi
if [ "${CONFIGS:-0}" = "1" ]; then
  for impl in ours reference; do
    timeout 900 $TR --master-port 29543 benchmark/bench_configs.py --impl $impl > gpurun_out/bench_configs_r2_${impl}_$N.jsonl 2> gpurun_out/bench_configs_r2_${impl}_$N.err
    grep '^{' gpurun_out/bench_configs_r2_${impl}_$N.jsonl | cut -c1-330
  done
fi

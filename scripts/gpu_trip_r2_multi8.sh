#!/bin/bash
# Round-2 8-GPU session: correctness at W=8, concurrency stress, fwd / fwd+bwd breakdown, headline bench (both
# arms, with the sampled oracle check and NVLink byte counters), BASELINE.json configs 2-5 (both arms).
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
export RFA_B200_PEER_TIMEOUT_S=30
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 300 -k "more_gpus and $N" > gpurun_out/pytest_multi_$N.log 2>&1; echo "multi tests exit $?"; tail -4 gpurun_out/pytest_multi_$N.log | cut -c1-200
STRESS_ITERS=1000 timeout 200 $TR --master-port 29547 benchmark/stress_overlap.py > gpurun_out/stress_overlap_$N.log 2>&1; echo "stress exit $?"; grep -E '^\{|differs|rfa:' gpurun_out/stress_overlap_$N.log | head -5
SWEEP=24 timeout 200 $TR --master-port 29542 benchmark/multi_breakdown.py > gpurun_out/breakdown_r2_$N.log 2>&1; grep -E "^t[0-9]" gpurun_out/breakdown_r2_$N.log | cut -c1-140
export RFA_B200_PEER_TIMEOUT_S=600
nvidia-smi nvlink -gt d -i 0 > gpurun_out/nvlink_before.txt 2>&1
timeout 300 $TR --master-port 29545 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_r2_ours_$N.log 2>&1; grep '"metric"' gpurun_out/bench_r2_ours_$N.log | cut -c1-1800
nvidia-smi nvlink -gt d -i 0 > gpurun_out/nvlink_after.txt 2>&1
timeout 300 $TR --master-port 29546 bench.py --impl reference --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_r2_reference_$N.log 2>&1; grep '"metric"' gpurun_out/bench_r2_reference_$N.log | cut -c1-700
for impl in ours reference; do
  timeout 400 $TR --master-port 29543 benchmark/bench_configs.py --impl $impl --steps 12 --warmup 5 > gpurun_out/bench_configs_r2_${impl}_$N.jsonl 2> gpurun_out/bench_configs_r2_${impl}_$N.err
  grep '^{' gpurun_out/bench_configs_r2_${impl}_$N.jsonl | cut -c1-330
done

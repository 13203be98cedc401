#!/bin/bash
# 8-GPU session: fused-path correctness, headline bench (both arms), reference-style iter/s tables.
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "gpus: $N"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== pytest"; timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 350 -k "more_gpus and $N" > gpurun_out/pytest_multi_$N.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_multi_$N.log; grep "rfa:" gpurun_out/pytest_multi_$N.log | sort | uniq -c | head -5
for impl in ours reference; do
  echo "== bench $impl N=$N"; timeout 400 $TR --master-port 29541 bench.py --gpus $N --steps 10 --warmup 5 --impl $impl > gpurun_out/bench_${impl}_$N.log 2>&1; echo "exit $?"; grep '"metric"' gpurun_out/bench_${impl}_$N.log | cut -c1-330
done
echo "== bench readme ours N=$N"; timeout 300 $TR --master-port 29543 bench.py --gpus $N --steps 10 --warmup 5 --config readme --no-e2e > gpurun_out/bench_readme_ours_$N.log 2>&1; grep '"metric"' gpurun_out/bench_readme_ours_$N.log | cut -c1-330
echo "== kvpacked fwd+bwd"; timeout 300 $TR --master-port 29545 benchmark/benchmark_kvpacked_func.py --num-iter 30 > gpurun_out/kvpacked_fwdbwd_$N.log 2>&1; grep "iter/s" gpurun_out/kvpacked_fwdbwd_$N.log
echo "== varlen fwd+bwd"; timeout 300 $TR --master-port 29547 benchmark/benchmark_varlen_kvpacked_func.py --num-iter 30 > gpurun_out/varlen_fwdbwd_$N.log 2>&1; grep "iter/s" gpurun_out/varlen_fwdbwd_$N.log; grep "rfa:" gpurun_out/varlen_fwdbwd_$N.log | sort | uniq -c | head -3

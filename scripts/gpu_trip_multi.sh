#!/bin/bash
# Multi-GPU session: correctness of the fused NVLink path + fallback, then bench at the available GPU count.
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "gpus: $N"
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
echo "== pytest multi"; timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 600 -x > gpurun_out/pytest_multi.log 2>&1; echo "pytest exit $?"; tail -30 gpurun_out/pytest_multi.log
for impl in reference ours; do
echo "== bench $impl N=$N"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 --impl $impl > gpurun_out/bench_${impl}_$N.log 2>&1; echo "exit $?"; tail -2 gpurun_out/bench_${impl}_$N.log | cut -c1-1500
done

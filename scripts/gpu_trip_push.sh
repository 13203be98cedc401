#!/bin/bash
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "== pytest fused (TMA push)"; timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 400 -k "2gpu and True" > gpurun_out/pytest_push.log 2>&1; echo "exit $?"; tail -3 gpurun_out/pytest_push.log
for tma in 1 0; do
  echo "== breakdown TMA=$tma"; RFA_B200_PUSH_TMA=$tma SWEEP=4,8,16,24 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 benchmark/multi_breakdown.py > gpurun_out/breakdown_tma$tma.log 2>&1; grep -E "^t[0-9]" gpurun_out/breakdown_tma$tma.log | head -8 | cut -c1-120
done

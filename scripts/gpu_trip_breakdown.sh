#!/bin/bash
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 benchmark/multi_breakdown.py > gpurun_out/breakdown_$N.log 2>&1; echo "exit $?"; grep -E "^t[0-9]" gpurun_out/breakdown_$N.log | tail -20; tail -5 gpurun_out/breakdown_$N.log

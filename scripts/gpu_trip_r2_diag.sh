#!/bin/bash
# Where does a fused step spend its time?  Same shards, three timings each: everything / compute only (no waits for
# remote K/V: wrong numbers, right amount of work) / communication only (K/V push + dK/dV return, no attention math).
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for mode in 0 2 4; do
  for ctas in ${PUSH_SWEEP:-24}; do
    echo "== flags=$mode push_ctas=$ctas"
    RFA_B200_FWD_FLAGS=$mode RFA_B200_BWD_FLAGS=$mode SWEEP=$ctas timeout 200 $TR --master-port 29542 benchmark/multi_breakdown.py > gpurun_out/diag_${N}_f${mode}_c${ctas}.log 2>&1
    grep -E "^t[0-9]" gpurun_out/diag_${N}_f${mode}_c${ctas}.log | cut -c1-140
  done
done

#!/bin/bash
# One ncu --set full capture of each hot kernel (1 GPU) + a per-launch duration list.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 8 -c 12 --csv --log-file gpurun_out/launches.csv python benchmark/ncu_target.py > gpurun_out/ncu_launches.log 2>&1
for k in attn_fwd_kernel attn_bwd_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/prof_$k python benchmark/ncu_target.py > gpurun_out/ncu_$k.log 2>&1
  echo "$k exit $?"
done
ls -la gpurun_out/*.ncu-rep

#!/bin/bash
# One ncu --set full capture of each hot kernel (1 GPU) + a per-launch duration list + N=1 bench (both arms).
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 8 -c 12 --csv --log-file gpurun_out/launches.csv python benchmark/ncu_target.py > gpurun_out/ncu_launches.log 2>&1
for k in attn_fwd_kernel attn_bwd_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/prof_$k python benchmark/ncu_target.py > gpurun_out/ncu_$k.log 2>&1
  echo "$k exit $?"
done
python benchmark/first_look.py > gpurun_out/first_look.log 2>&1; tail -3 gpurun_out/first_look.log | cut -c1-200
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_ours_1.log 2>&1; grep '"metric"' gpurun_out/bench_ours_1.log | cut -c1-300
python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_reference_1.log 2>&1; grep '"metric"' gpurun_out/bench_reference_1.log | cut -c1-300

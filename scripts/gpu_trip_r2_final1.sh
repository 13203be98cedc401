#!/bin/bash
# Final single-GPU session of round 2: exactly the driver's GPU tier + smoke(), the launch list of one fwd+bwd step
# (which kernels run: evidence for "no ATen glue"), one ncu --set full capture per hot kernel, bench.py for both arms.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/ -x -q -m gpu --timeout 600 > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest -m gpu exit $?"; tail -4 gpurun_out/pytest_gpu_final.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke_final.log
ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 12 --csv --log-file gpurun_out/launches_1gpu_r2.csv python benchmark/ncu_target.py > gpurun_out/ncu_launches.log 2>&1; echo "launch list exit $?"
for k in attn_fwd_kernel attn_bwd_kernel; do
  ITERS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/prof_r2_$k python benchmark/ncu_target.py > gpurun_out/ncu_r2_$k.log 2>&1
  echo "ncu $k exit $?"
done
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_final_ours_1.log 2>&1; grep '"metric"' gpurun_out/bench_final_ours_1.log | cut -c1-600
timeout 300 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_final_reference_1.log 2>&1; grep '"metric"' gpurun_out/bench_final_reference_1.log | cut -c1-400

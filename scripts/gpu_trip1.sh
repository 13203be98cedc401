#!/bin/bash
# First GPU session: descriptor probe, kernel tests, kernel timings vs flash_attn, bench (both arms).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
echo "== probe"; timeout 300 python benchmark/probe_descriptors.py > gpurun_out/probe.log 2>&1; echo "probe exit $?"; tail -12 gpurun_out/probe.log
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x > gpurun_out/pytest.log 2>&1; echo "pytest exit $?"; tail -25 gpurun_out/pytest.log
echo "== first look"; timeout 600 python benchmark/first_look.py > gpurun_out/first_look.log 2>&1; echo "first_look exit $?"; tail -8 gpurun_out/first_look.log
echo "== bench ref"; timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref.log 2>&1; echo "exit $?"; tail -3 gpurun_out/bench_ref.log
echo "== bench ours"; timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_ours.log 2>&1; echo "exit $?"; tail -3 gpurun_out/bench_ours.log

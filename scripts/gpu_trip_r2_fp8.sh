#!/bin/bash
# Block-scaled fp8 forward: kernel tests (1 GPU part runs on GPU 0), fused 2-GPU test when 2 GPUs are visible, and the
# stripe fp8 config forward-only: dequantise path vs native kernel.
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
export RFA_B200_PEER_TIMEOUT_S=30
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 300 -k "fp8" > gpurun_out/pytest_fp8.log 2>&1; echo "fp8 kernel tests exit $?"; tail -5 gpurun_out/pytest_fp8.log | cut -c1-300
if [ "$N" -ge 2 ]; then
  timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 300 -k "fp8 or batch_schemes_2gpu" > gpurun_out/pytest_fp8_multi.log 2>&1; echo "fp8 fused tests exit $?"; tail -5 gpurun_out/pytest_fp8_multi.log | cut -c1-300
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
  for lvl in 0 2; do
    RFA_B200_FP8_KERNEL=$lvl timeout 300 $TR --master-port 29544 benchmark/bench_configs.py --only stripe8 --forward-only > gpurun_out/bench_stripe8_blockscaled_fp8kernel$lvl.jsonl 2> gpurun_out/bench_stripe8_blockscaled_fp8kernel$lvl.err
    grep '^{' gpurun_out/bench_stripe8_blockscaled_fp8kernel$lvl.jsonl | cut -c1-260
  done
fi

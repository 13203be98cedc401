#!/bin/bash
# K/V transport experiments (N = all visible GPUs): in-kernel TMA push (completion published per destination) vs
# copy engines on a side stream; correctness first, then everything / compute-only / communication-only timings.
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
export RFA_B200_PEER_TIMEOUT_S=30
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for tr in ${TRANSPORTS:-push dma}; do
  if [ "${TESTS:-1}" = "1" ]; then
    RFA_B200_KV_TRANSPORT=$tr timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 300 -k "${PYTEST_K:-batch_schemes_2gpu and True}" > gpurun_out/pytest_transport_${tr}_$N.log 2>&1; echo "transport=$tr tests exit $?"; tail -3 gpurun_out/pytest_transport_${tr}_$N.log | cut -c1-200
  fi
  for mode in ${MODES:-0 4}; do
    echo "== transport=$tr flags=$mode"
    RFA_B200_KV_TRANSPORT=$tr RFA_B200_FWD_FLAGS=$mode RFA_B200_BWD_FLAGS=$mode SWEEP=${PUSH_SWEEP:-24} timeout 200 $TR --master-port 29542 benchmark/multi_breakdown.py > gpurun_out/transport_${N}_${tr}_f${mode}.log 2>&1
    grep -E "^t[0-9]|rror" gpurun_out/transport_${N}_${tr}_f${mode}.log | cut -c1-140
  done
  if [ "${BENCH:-0}" = "1" ]; then
    RFA_B200_PEER_TIMEOUT_S=600 RFA_B200_KV_TRANSPORT=$tr timeout 300 $TR --master-port 29545 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_r2_${tr}_$N.log 2>&1; grep '"metric"' gpurun_out/bench_r2_${tr}_$N.log | cut -c1-900
  fi
done

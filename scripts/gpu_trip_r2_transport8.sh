#!/bin/bash
# 8-GPU decision run: full-step time of the headline shard for three K/V transports, communication-only time of the
# in-kernel push, then bench.py (with the oracle check) for the fastest one.
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
export RFA_B200_PEER_TIMEOUT_S=60
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
best=""; best_ms=1000000
for cfg in "push 24" "push 48" "dma 24"; do
  set -- $cfg
  RFA_B200_KV_TRANSPORT=$1 SWEEP=$2 timeout 200 $TR --master-port 29542 benchmark/multi_breakdown.py > gpurun_out/transport_${N}_$1_$2.log 2>&1
  line=$(grep -E "^t4096" gpurun_out/transport_${N}_$1_$2.log | head -1)
  echo "transport=$1 push_ctas=$2: $line $(grep -E '^t8192' gpurun_out/transport_${N}_$1_$2.log | head -1 | cut -c1-110)"
  ms=$(echo "$line" | sed -n "s/.*'fwdbwd_ms': \([0-9.]*\).*/\1/p")
  if [ -n "$ms" ] && python -c "import sys; sys.exit(0 if float('$ms') < float('$best_ms') else 1)"; then best_ms=$ms; best="$1 $2"; fi
done
RFA_B200_KV_TRANSPORT=push RFA_B200_FWD_FLAGS=4 RFA_B200_BWD_FLAGS=4 SWEEP=24 timeout 200 $TR --master-port 29542 benchmark/multi_breakdown.py > gpurun_out/transport_${N}_push_24_commonly.log 2>&1
echo "comm-only push 24: $(grep -E '^t4096' gpurun_out/transport_${N}_push_24_commonly.log | head -1)"
set -- $best
echo "== bench.py with transport=$1 push_ctas=$2 (fastest full step: $best_ms ms)"
RFA_B200_PEER_TIMEOUT_S=600 RFA_B200_KV_TRANSPORT=$1 RFA_B200_PUSH_CTAS=$2 timeout 300 $TR --master-port 29545 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_r2_best_$N.log 2>&1; grep '"metric"' gpurun_out/bench_r2_best_$N.log | cut -c1-1700

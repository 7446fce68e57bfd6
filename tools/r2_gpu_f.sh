#!/bin/bash
# round 2, GPU call F (N GPUs = $1): gather throughput against NCCL's point-to-point channel count
N=${1:-2}
mkdir -p gpurun_out
for ch in default 2 4 8; do
  if [ "$ch" = default ]; then unset NCCL_MAX_P2P_NCHANNELS; else export NCCL_MAX_P2P_NCHANNELS=$ch; fi
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29714 bench.py --gpus $N --quick --no-cpu --steps 10 --warmup 3 > gpurun_out/r2f_n${N}_ch$ch.json 2> gpurun_out/r2f_n${N}_ch$ch.err
  echo "ch=$ch"; cat gpurun_out/r2f_n${N}_ch$ch.json
done

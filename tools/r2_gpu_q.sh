#!/bin/bash
# round 2, GPU call Q (N GPUs = $1): gather legs -- frames per wire operation, transport, trace streams (bench.py --quick)
N=${1:-8}
mkdir -p gpurun_out
export ORT_BENCH_GATHER_LEGS="round_robin:12:1:8,round_robin:8:1:8,round_robin:48:1:8,round_robin:24:1:4,round_robin:12:0:8,round_robin:6:0:8,round_robin:24:1:8"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29731 bench.py --gpus $N --quick --no-cpu --steps 10 --warmup 3 > gpurun_out/r2q_n${N}_quick.json 2> gpurun_out/r2q_n${N}_quick.err
cat gpurun_out/r2q_n${N}_quick.json; tail -3 gpurun_out/r2q_n${N}_quick.err

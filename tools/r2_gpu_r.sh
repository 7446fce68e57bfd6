#!/bin/bash
# round 2, GPU call R (N GPUs = $1): quick bench line at N with the gather legs (N = 4: frames per wire operation 6 / 12)
N=${1:-4}
mkdir -p gpurun_out
export ORT_BENCH_GATHER_LEGS="round_robin:6:1:4,round_robin:12:1:4,round_robin:12:1:8,round_robin:12:0:4"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29741 bench.py --gpus $N --quick --no-cpu --steps 10 --warmup 3 > gpurun_out/r2r_n${N}_quick.json 2> gpurun_out/r2r_n${N}_quick.err
cat gpurun_out/r2r_n${N}_quick.json; tail -3 gpurun_out/r2r_n${N}_quick.err

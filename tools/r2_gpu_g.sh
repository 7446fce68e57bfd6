#!/bin/bash
# round 2, GPU call G (8 GPUs): weak-scaling value + gather legs (quick mode), default NCCL settings
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29715 bench.py --gpus 8 --quick --no-cpu --steps 10 --warmup 3 > gpurun_out/r2g_n8_quick.json 2> gpurun_out/r2g_n8_quick.err
cat gpurun_out/r2g_n8_quick.json; tail -3 gpurun_out/r2g_n8_quick.err

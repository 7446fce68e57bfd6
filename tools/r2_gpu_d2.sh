#!/bin/bash
# round 2, GPU call D2 (2 GPUs): the library's NCCL gather on real devices -- parity test, then the N=2 bench line
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_configs.py -m gpu -x -q -k "multi_gpu" ) > gpurun_out/r2d2_mgtest.log 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2d2_bench_n2.json 2> gpurun_out/r2d2_bench_n2.err
tail -15 gpurun_out/r2d2_mgtest.log; cat gpurun_out/r2d2_bench_n2.json; tail -5 gpurun_out/r2d2_bench_n2.err

#!/bin/bash
# round 2, GPU call H (N GPUs = $1): peer-copy transport -- hardware parity test (both transports), then quick bench with gather legs
N=${1:-2}
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_configs.py -m gpu -x -q -k "multi_gpu" ) > gpurun_out/r2h_mgtest_n$N.log 2>&1
tail -12 gpurun_out/r2h_mgtest_n$N.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29716 bench.py --gpus $N --quick --no-cpu --steps 10 --warmup 3 > gpurun_out/r2h_n${N}_quick.json 2> gpurun_out/r2h_n${N}_quick.err
cat gpurun_out/r2h_n${N}_quick.json; tail -4 gpurun_out/r2h_n${N}_quick.err

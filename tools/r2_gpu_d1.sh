#!/bin/bash
# round 2, GPU call D1 (1 GPU): LeanWalker after the advance rewrite -- loop shapes / register budgets of the product kernel
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2d1_tests.log
for sh in 0 1 2 3; do
  python bench.py --quick --no-cpu --steps 10 --warmup 3 --opt lean_shape=$sh > gpurun_out/r2d1_shape$sh.json 2> gpurun_out/r2d1_shape$sh.err
done
python bench.py --quick --no-cpu --steps 10 --warmup 3 --opt lean_shape=0 --opt l1_carveout=30 > gpurun_out/r2d1_shape0_co30.json 2>> gpurun_out/r2d1_shape0.err
cat gpurun_out/r2d1_tests.log gpurun_out/r2d1_shape*.json

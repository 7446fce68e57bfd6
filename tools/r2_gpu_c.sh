#!/bin/bash
# round 2, GPU call C: GPU suite, the new bench line (N=1) + reference arm, loop-shape / stack / register experiments
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q --durations=6 ) > gpurun_out/r2c_tests.log 2>&1
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2c_ref.json 2> gpurun_out/r2c_ref.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r2c_bench_n1.json 2> gpurun_out/r2c_bench_n1.err
for v in 16 17 18 19; do
  ORT_B200_EXPERIMENTS=1 python bench.py --quick --no-cpu --steps 10 --warmup 3 --variant $v > gpurun_out/r2c_quick_v$v.json 2> gpurun_out/r2c_quick_v$v.err
done
tail -12 gpurun_out/r2c_tests.log; cat gpurun_out/r2c_ref.json; cat gpurun_out/r2c_bench_n1.json; tail -3 gpurun_out/r2c_bench_n1.err; cat gpurun_out/r2c_quick_v*.json

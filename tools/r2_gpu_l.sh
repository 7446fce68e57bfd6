#!/bin/bash
# round 2, GPU call L (1 GPU): beam start with whole-tile MISSes -- GPU suite, full N=1 bench line, quick A/B, ncu of one step
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q --durations=5 ) > gpurun_out/r2l_tests.log 2>&1
tail -12 gpurun_out/r2l_tests.log
for b in 0 1; do
  python bench.py --quick --no-cpu --steps 10 --warmup 3 --opt beam=$b > gpurun_out/r2l_quick_beam$b.json 2> gpurun_out/r2l_quick_beam$b.err
  cat gpurun_out/r2l_quick_beam$b.json
done
python bench.py --steps 20 --warmup 5 > gpurun_out/r2l_bench_n1.json 2> gpurun_out/r2l_bench_n1.err
cat gpurun_out/r2l_bench_n1.json; tail -2 gpurun_out/r2l_bench_n1.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"trace_frame_kernel|beam_start_kernel" --launch-skip 27 -c 6 -f -o gpurun_out/r2l_beam \
  python bench.py --quick --no-cpu --steps 1 --warmup 3 > gpurun_out/r2l_ncu_full.log 2>&1
tail -3 gpurun_out/r2l_ncu_full.log

#!/bin/bash
# round 2, GPU call B: whole GPU suite (new full-size tests), lean vs round-1 walker, experiment variants 14/15, config 3, ncu
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q --durations=8 ) > gpurun_out/r2b_tests.log 2>&1
for v in 1 13; do
  python bench.py --quick --steps 10 --warmup 3 --variant $v > gpurun_out/r2b_quick_v$v.json 2> gpurun_out/r2b_quick_v$v.err
done
for v in 14 15; do
  ORT_B200_EXPERIMENTS=1 python bench.py --quick --steps 10 --warmup 3 --variant $v > gpurun_out/r2b_quick_v$v.json 2> gpurun_out/r2b_quick_v$v.err
done
python tools/bench_rays.py > gpurun_out/r2b_rays.json 2> gpurun_out/r2b_rays.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_frame_kernel --launch-skip 3 -c 3 -f -o gpurun_out/r2b_lean \
  python bench.py --quick --steps 1 --warmup 3 --variant 13 > gpurun_out/r2b_ncu13.log 2>&1
for v in 14 15; do
  ORT_B200_EXPERIMENTS=1 timeout 600 ncu --set full --clock-control none -k regex:trace_frame_walker --launch-skip 3 -c 3 -f -o gpurun_out/r2b_v$v \
    python bench.py --quick --steps 1 --warmup 3 --variant $v > gpurun_out/r2b_ncu$v.log 2>&1
done
tail -15 gpurun_out/r2b_tests.log; cat gpurun_out/r2b_quick_v*.json; tail -5 gpurun_out/r2b_rays.err

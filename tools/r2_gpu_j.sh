#!/bin/bash
# round 2, GPU call J (1 GPU): beam start -- its GPU tests, then the whole GPU suite, quick bench with the option off / on,
# forced coarser grid levels, and an ncu --set full of the step's frame launches + marches
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_beam.py -m gpu -x -q --durations=5 ) > gpurun_out/r2j_beam_tests.log 2>&1
tail -25 gpurun_out/r2j_beam_tests.log
for b in 0 1; do
  python bench.py --quick --no-cpu --steps 10 --warmup 3 --opt beam=$b > gpurun_out/r2j_quick_beam$b.json 2> gpurun_out/r2j_quick_beam$b.err
  cat gpurun_out/r2j_quick_beam$b.json; tail -2 gpurun_out/r2j_quick_beam$b.err
done
for lv in 5 6; do
  python bench.py --quick --no-cpu --steps 10 --warmup 3 --opt beam_level=$lv > gpurun_out/r2j_quick_level$lv.json 2> gpurun_out/r2j_quick_level$lv.err
  cat gpurun_out/r2j_quick_level$lv.json
done
( time python -m pytest tests -m gpu -x -q --durations=5 ) > gpurun_out/r2j_tests.log 2>&1
tail -12 gpurun_out/r2j_tests.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"trace_frame_kernel|beam_start_kernel" --launch-skip 6 -c 6 -f -o gpurun_out/r2j_beam \
  python bench.py --quick --no-cpu --steps 1 --warmup 3 > gpurun_out/r2j_ncu_full.log 2>&1
tail -3 gpurun_out/r2j_ncu_full.log

#!/bin/bash
# round 2, GPU call A: parity tests, lean walker (variant 13) vs round-1 default (variant 1), ncu of the lean kernel
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2a_tests.log
for v in 1 13; do
  python bench.py --quick --steps 10 --warmup 3 --variant $v > gpurun_out/r2a_quick_v$v.json 2> gpurun_out/r2a_quick_v$v.err
done
python bench.py --quick --steps 10 --warmup 3 --variant 13 --opt l1_carveout=50 > gpurun_out/r2a_quick_v13_co50.json 2>> gpurun_out/r2a_quick_v13.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_frame_lean --launch-skip 3 -c 3 -f -o gpurun_out/r2a_lean \
  python bench.py --quick --steps 1 --warmup 3 --variant 13 > gpurun_out/r2a_ncu.log 2>&1
cat gpurun_out/r2a_tests.log gpurun_out/r2a_quick_v1.json gpurun_out/r2a_quick_v13.json gpurun_out/r2a_quick_v13_co50.json

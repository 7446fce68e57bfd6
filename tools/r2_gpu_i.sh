#!/bin/bash
# round 2, GPU call I (1 GPU): state of HEAD after the container was re-created -- whole GPU suite, N=1 bench line + reference arm,
# ncu launch list of the bench command, ncu --set full of the step's three frame launches (refreshes profiles/traffic.json)
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q --durations=6 ) > gpurun_out/r2i_tests.log 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2i_ref.json 2> gpurun_out/r2i_ref.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r2i_bench_n1.json 2> gpurun_out/r2i_bench_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2i_launches.csv \
  python bench.py --quick --no-cpu --steps 2 --warmup 3 > gpurun_out/r2i_ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_frame_kernel --launch-skip 3 -c 3 -f -o gpurun_out/r2i_lean \
  python bench.py --quick --no-cpu --steps 1 --warmup 3 > gpurun_out/r2i_ncu_full.log 2>&1
tail -12 gpurun_out/r2i_tests.log; cat gpurun_out/r2i_ref.json; cat gpurun_out/r2i_bench_n1.json; tail -3 gpurun_out/r2i_bench_n1.err; tail -3 gpurun_out/r2i_ncu_full.log

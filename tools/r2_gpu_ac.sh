#!/bin/bash
# round 2, GPU call AC (1 GPU): ncu --set full of strip launches (the share of rank 3 of 8) of the final build, next to the
# whole-frame capture profiles/r2_final_beam_ncu_full.md
mkdir -p gpurun_out
timeout 500 ncu --set full --clock-control none -k regex:"trace_frame_kernel|beam_start_kernel" --launch-skip 157 -c 6 -f -o gpurun_out/r2ac_strips \
  python bench.py --quick --no-cpu --steps 1 --warmup 3 --as-rank 3/8 > gpurun_out/r2ac_ncu.log 2>&1
tail -2 gpurun_out/r2ac_ncu.log

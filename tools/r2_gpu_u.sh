#!/bin/bash
# round 2, GPU call U (1 GPU): walkers / loop shapes WITH the beam start (experiment variants 24-27) against the product
mkdir -p gpurun_out
export ORT_B200_EXPERIMENTS=1
python -m pytest tests/test_gpu_beam.py -m gpu -x -q -k beam_experiment_walkers 2>&1 | tail -3
for v in 13 24 25 26 27; do
  python bench.py --quick --no-cpu --steps 20 --warmup 5 --variant $v --opt band_order=0 > gpurun_out/r2u_quick_v$v.json 2> gpurun_out/r2u_quick_v$v.err
  python - <<PY
import json; d=json.load(open('gpurun_out/r2u_quick_v$v.json')); print("variant $v:", d["value"], d["serial_value"], d["per_frame_ms_serial"], d["beam_levels"])
PY
  tail -1 gpurun_out/r2u_quick_v$v.err | cut -c1-200
done
timeout 600 ncu --set full --clock-control none -k regex:"trace_frame_walker_beam_kernel" --launch-skip 12 -c 3 -f -o gpurun_out/r2u_v24 \
  python bench.py --quick --no-cpu --steps 1 --warmup 3 --variant 24 --opt band_order=0 > gpurun_out/r2u_ncu24.log 2>&1
tail -2 gpurun_out/r2u_ncu24.log

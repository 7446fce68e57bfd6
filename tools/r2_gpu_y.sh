#!/bin/bash
# round 2, GPU call Y (1 GPU): what the beam guard and the register budget cost (experiment variants 27-31)
mkdir -p gpurun_out
export ORT_B200_EXPERIMENTS=1
for v in 13 27 28 29 30 31; do
  python bench.py --quick --no-cpu --steps 20 --warmup 5 --variant $v > gpurun_out/r2y_quick_v$v.json 2> gpurun_out/r2y_quick_v$v.err
  python - <<PY
import json; d=json.load(open('gpurun_out/r2y_quick_v$v.json')); print("variant $v:", d["value"], d["serial_value"], d["per_frame_ms_serial"])
PY
done

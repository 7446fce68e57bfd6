#!/bin/bash
# round 2, GPU call S (1 GPU): the share of one rank of 8 / of 4 -- tile heights and stream counts with the beam start
mkdir -p gpurun_out
for tr in 8 16 32; do
  python bench.py --quick --no-cpu --steps 20 --warmup 5 --as-rank 3/8 --tile-rows $tr > gpurun_out/r2s_asrank8_tr$tr.json 2> gpurun_out/r2s_asrank8_tr$tr.err
  python - <<PY
import json; d=json.load(open('gpurun_out/r2s_asrank8_tr$tr.json')); print("3/8 tile_rows $tr:", d["value"], d["ms_per_step"], d["streams"], d["host_enqueue_ms_per_step"])
PY
done
for st in 4 6 8; do
  python bench.py --quick --no-cpu --steps 20 --warmup 5 --as-rank 1/4 --streams $st > gpurun_out/r2s_asrank4_s$st.json 2> gpurun_out/r2s_asrank4_s$st.err
  python - <<PY
import json; d=json.load(open('gpurun_out/r2s_asrank4_s$st.json')); print("1/4 streams $st:", d["value"], d["ms_per_step"], d["host_enqueue_ms_per_step"])
PY
done
for st in 3 4 6; do
  python bench.py --quick --no-cpu --steps 20 --warmup 5 --streams $st > gpurun_out/r2s_n1_s$st.json 2> gpurun_out/r2s_n1_s$st.err
  python - <<PY
import json; d=json.load(open('gpurun_out/r2s_n1_s$st.json')); print("N=1 streams $st:", d["value"], d["ms_per_step"])
PY
done

#!/bin/bash
# round 2, GPU call P (1 GPU): the final build -- whole GPU suite, reference arm, N=1 bench line, ncu launch list of the bench
# command, ncu --set full of one step's frame launches + marches (-> profiles/traffic.json with the final kernel-source hash)
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q --durations=6 ) > gpurun_out/r2p_tests.log 2>&1
tail -12 gpurun_out/r2p_tests.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2p_ref.json 2> gpurun_out/r2p_ref.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r2p_bench_n1.json 2> gpurun_out/r2p_bench_n1.err
tail -2 gpurun_out/r2p_bench_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2p_launches.csv \
  python bench.py --quick --no-cpu --steps 2 --warmup 3 > gpurun_out/r2p_ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"trace_frame_kernel|beam_start_kernel" --launch-skip 27 -c 6 -f -o gpurun_out/r2p_final \
  python bench.py --quick --no-cpu --steps 1 --warmup 3 > gpurun_out/r2p_ncu_full.log 2>&1
tail -2 gpurun_out/r2p_ncu_full.log
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2p_bench_n1.json')); r=json.load(open('gpurun_out/r2p_ref.json'))
print("value",d["value"],"serial",d["serial"]["value"],d["serial"]["per_launch_ms"],"e2e",d["e2e"]["value"],"rgba",d["e2e_rgba"]["value"],"ref",r["value"],"parity",d["parity"])
PY

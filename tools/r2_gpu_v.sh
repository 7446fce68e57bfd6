#!/bin/bash
# round 2, GPU call V (1 GPU): configs 4 and 5 with the beam start off / on (config 4: the grid is rebuilt after every edit)
mkdir -p gpurun_out
for b in 0 1; do
  python tools/bench_configs.py --config 4 --bulk --opt beam=$b > gpurun_out/r2v_config4_beam$b.json 2> gpurun_out/r2v_config4_beam$b.err; cat gpurun_out/r2v_config4_beam$b.json; tail -1 gpurun_out/r2v_config4_beam$b.err | cut -c1-300
  python tools/bench_configs.py --config 4 --bulk --depth 12 --opt beam=$b > gpurun_out/r2v_config4_d12_beam$b.json 2> gpurun_out/r2v_config4_d12_beam$b.err; cat gpurun_out/r2v_config4_d12_beam$b.json
done
for b in 0 1; do
  python tools/bench_configs.py --config 5 --opt beam=$b > gpurun_out/r2v_config5_beam$b.json 2> gpurun_out/r2v_config5_beam$b.err; cat gpurun_out/r2v_config5_beam$b.json; tail -1 gpurun_out/r2v_config5_beam$b.err | cut -c1-300
done

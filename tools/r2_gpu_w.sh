#!/bin/bash
# round 2, GPU call W (1 GPU): lazy beam grids ("beam_after") -- GPU suite, config 4 with edits / steady, N=1 quick bench
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q --durations=4 ) > gpurun_out/r2w_tests.log 2>&1
tail -9 gpurun_out/r2w_tests.log
python tools/bench_configs.py --config 4 --bulk > gpurun_out/r2w_config4_default.json 2> gpurun_out/r2w_config4_default.err; cat gpurun_out/r2w_config4_default.json
python tools/bench_configs.py --config 4 --bulk --depth 12 > gpurun_out/r2w_config4_d12_default.json 2>> gpurun_out/r2w_config4_default.err; cat gpurun_out/r2w_config4_d12_default.json
for b in 0 1; do
  python tools/bench_configs.py --config 4 --bulk --no-edits --opt beam=$b > gpurun_out/r2w_config4_steady_beam$b.json 2>> gpurun_out/r2w_config4_default.err; cat gpurun_out/r2w_config4_steady_beam$b.json
done
python bench.py --quick --no-cpu --steps 20 --warmup 5 > gpurun_out/r2w_quick.json 2> gpurun_out/r2w_quick.err; cat gpurun_out/r2w_quick.json
python -c "import __graft_entry__ as g; g.smoke()"

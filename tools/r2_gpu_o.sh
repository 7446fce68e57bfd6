#!/bin/bash
# round 2, GPU call O (1 GPU): band schedule with the recording kernel split off; the share of one rank of 8 -- streams vs one
# batched launch per step, with the host's enqueue time beside the GPU time
mkdir -p gpurun_out
python -m pytest tests/test_gpu_beam.py -m gpu -x -q 2>&1 | tail -3
for b in 0 1; do
  python bench.py --quick --no-cpu --steps 20 --warmup 5 --opt band_order=$b > gpurun_out/r2o_quick_band$b.json 2> gpurun_out/r2o_quick_band$b.err
  cat gpurun_out/r2o_quick_band$b.json; tail -1 gpurun_out/r2o_quick_band$b.err
done
for mode in streams batch; do
  python bench.py --quick --no-cpu --steps 20 --warmup 5 --as-rank 3/8 --launch $mode > gpurun_out/r2o_asrank_$mode.json 2> gpurun_out/r2o_asrank_$mode.err
  cat gpurun_out/r2o_asrank_$mode.json; tail -1 gpurun_out/r2o_asrank_$mode.err
done
python bench.py --quick --no-cpu --steps 20 --warmup 5 --as-rank 3/8 --streams 12 > gpurun_out/r2o_asrank_s12.json 2> gpurun_out/r2o_asrank_s12.err
cat gpurun_out/r2o_asrank_s12.json
python bench.py --quick --no-cpu --steps 20 --warmup 5 --as-rank 1/2 > gpurun_out/r2o_asrank_1of2.json 2> gpurun_out/r2o_asrank_1of2.err
cat gpurun_out/r2o_asrank_1of2.json

#!/bin/bash
# round 2, GPU call N (1 GPU): band schedule from recorded costs -- beam tests, quick bench with the option off / on, and the
# share of one rank of 8 (--as-rank) off / on
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_beam.py -m gpu -x -q ) > gpurun_out/r2n_beam_tests.log 2>&1
tail -8 gpurun_out/r2n_beam_tests.log
for b in 0 1; do
  python bench.py --quick --no-cpu --steps 20 --warmup 5 --opt band_order=$b > gpurun_out/r2n_quick_band$b.json 2> gpurun_out/r2n_quick_band$b.err
  cat gpurun_out/r2n_quick_band$b.json; tail -1 gpurun_out/r2n_quick_band$b.err
  python bench.py --quick --no-cpu --steps 20 --warmup 5 --as-rank 3/8 --opt band_order=$b > gpurun_out/r2n_asrank_band$b.json 2> gpurun_out/r2n_asrank_band$b.err
  cat gpurun_out/r2n_asrank_band$b.json
done
python bench.py --quick --no-cpu --steps 20 --warmup 5 --as-rank 3/8 --opt beam=0 --opt band_order=0 > gpurun_out/r2n_asrank_nobeam.json 2> gpurun_out/r2n_asrank_nobeam.err
cat gpurun_out/r2n_asrank_nobeam.json

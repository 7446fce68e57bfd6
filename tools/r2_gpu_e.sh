#!/bin/bash
# round 2, GPU call E (2 GPUs): leaner prologue + multi-stream strip tracing inside ort_mg; tests on GPU 0, bench on 2
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r2e_tests.log 2>&1
CUDA_VISIBLE_DEVICES=0 python bench.py --quick --no-cpu --steps 10 --warmup 3 > gpurun_out/r2e_quick_n1.json 2> gpurun_out/r2e_quick_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29713 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2e_bench_n2.json 2> gpurun_out/r2e_bench_n2.err
tail -6 gpurun_out/r2e_tests.log; cat gpurun_out/r2e_quick_n1.json; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2e_bench_n2.json'))
print("N=2 value",d["value"],"e2e",d["e2e"]["value"]); print(json.dumps(d["with_gather"],indent=1))
PY
tail -3 gpurun_out/r2e_bench_n2.err

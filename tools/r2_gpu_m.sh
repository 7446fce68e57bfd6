#!/bin/bash
# round 2, GPU call M (N GPUs = $1): beam start under ort_mg -- hardware multi-GPU parity test, then the full bench line at N
N=${1:-2}
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_configs.py -m gpu -x -q -k "multi_gpu" ) > gpurun_out/r2m_mgtest_n$N.log 2>&1
tail -6 gpurun_out/r2m_mgtest_n$N.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2m_bench_n$N.json 2> gpurun_out/r2m_bench_n$N.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2m_bench_n$N.json'))
print("N=$N value",d["value"],"no_gather",d.get("no_gather",{}).get("value"),"e2e",d["e2e"]["value"],"parity",d["parity"])
print(json.dumps(d.get("with_gather"),indent=1)[:3000])
print(d["roofline"]["frac"], d["roofline"]["simt"])
PY
tail -3 gpurun_out/r2m_bench_n$N.err

#!/bin/bash
# round 2, GPU call AB (N GPUs = $1): end-to-end legs -- launches per host-buffer frame (1 / 2 / 4) for the strips of N ranks
N=${1:-8}
mkdir -p gpurun_out
export ORT_BENCH_E2E_CHUNKS="1,2,4"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29751 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2ab_bench_n$N.json 2> gpurun_out/r2ab_bench_n$N.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2ab_bench_n$N.json'))
print("N=$N value",d["value"],"no_gather",d["no_gather"]["value"],"e2e",d["e2e"]["value"],d["e2e"].get("chunk_legs_Mrays/s"),"ingest",d["e2e"]["host_ingest_peak_gbs"],"rgba",d["e2e_rgba"]["value"])
PY
tail -2 gpurun_out/r2ab_bench_n$N.err

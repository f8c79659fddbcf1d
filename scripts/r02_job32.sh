#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n${N}_pipe.json 2> gpurun_out/bench_n${N}_pipe.err
echo "== N=$N rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_n${N}_pipe.json").read().strip().splitlines()[-1])
    print(d["ms_per_step"], d["config"]["isolated_ms_per_msm"], d["stages_ms"], d.get("bucket_reduce_scatter",{}).get("last_exchange_us"), {k:v for k,v in (d.get("e2e") or {}).items() if k in ("ms_per_step","pageable")})
except Exception as ex:
    print("no json:", ex)
PY
tail -3 gpurun_out/bench_n${N}_pipe.err | cut -c1-300

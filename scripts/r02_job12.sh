#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_job12.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1_job12.json 2> gpurun_out/bench_n1_job12.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n1_job12.json").read().strip().splitlines()[-1])
for k in ("ms_per_step","stages_ms","int_pipe","e2e","contest_call","one_call","sweep","bls12_381_2pow24","fastest_entry"):
    print(k, json.dumps(d.get(k))[:900])
PY
tail -3 gpurun_out/bench_n1_job12.err

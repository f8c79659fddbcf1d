#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_job26.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1_job26.json 2> gpurun_out/bench_n1_job26.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n1_job26.json").read().strip().splitlines()[-1])
for k in ("ms_per_step","stages_ms","int_pipe","roofline","e2e","contest_call","one_call","cpu_baseline","fastest_entry"):
    print(k, json.dumps(d.get(k))[:700])
for r in d.get("sweep") or []: print("sweep", r)
print("381", d.get("bls12_381_2pow24"))
PY
tail -3 gpurun_out/bench_n1_job26.err

#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_job31.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-sweep > gpurun_out/bench_n1_job31.json 2> gpurun_out/bench_n1_job31.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n1_job31.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["config"]["isolated_ms_per_msm"], d["stages_ms"], d["e2e"]["ms_per_step"], d["contest_call"]["pageable_ms_total"], d["one_call"]["ms"])
PY
tail -3 gpurun_out/bench_n1_job31.err

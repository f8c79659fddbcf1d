#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none -s 64 -c 60 --csv --log-file gpurun_out/r02_launches26.csv python scripts/sweep.py 0:26 > /dev/null 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1_final.json 2> gpurun_out/bench_n1_final.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n1_final.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["stages_ms"], d["e2e"]["ms_per_step"], d["contest_call"]["pageable_ms_total"], d["one_call"]["ms"], [r.get("bit_exact") for r in d["sweep"]], d["bls12_381_2pow24"].get("bit_exact"))
PY

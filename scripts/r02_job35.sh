#!/bin/bash
mkdir -p gpurun_out
( time python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_n1_driverlike.json 2> gpurun_out/bench_n1_driverlike.err ) 2>&1 | grep real
wc -l gpurun_out/bench_n1_driverlike.json
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n1_driverlike.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["config"]["isolated_ms_per_msm"], d["e2e"]["ms_per_step"], d["contest_call"]["pageable_ms_total"], d["one_call"]["ms"], [ (r["log_n"], r.get("ms"), r.get("bit_exact")) for r in d["sweep"]], d["bls12_381_2pow24"].get("ms"), d["fastest_entry"]["ratio"], d["gpu_launches"], d["clocks"])
PY
( time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ref_driverlike.json 2> gpurun_out/bench_ref_driverlike.err ) 2>&1 | grep real
cut -c1-600 gpurun_out/bench_ref_driverlike.json

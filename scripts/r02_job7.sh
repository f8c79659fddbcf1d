#!/bin/bash
# 2 GPUs: the bucket reduce-scatter between processes (CUDA IPC) against the all-gather-only path
mkdir -p gpurun_out
N=${1:-2}
for mode in "" "--no-peer-reduce"; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 $mode \
     > gpurun_out/bench_n${N}${mode}.json 2> gpurun_out/bench_n${N}${mode}.err
  echo "== N=$N $mode rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_n${N}${mode}.json").read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("ms_per_step","stages_ms","bucket_reduce_scatter")}, {k:d["e2e"][k] for k in ("ms_per_step","pinned","api")} if d.get("e2e") else None)
except Exception as ex:
    print("no json:", ex)
PY
  grep -c "NCCL INFO" gpurun_out/bench_n${N}${mode}.err; grep -m2 "nranks" gpurun_out/bench_n${N}${mode}.err | cut -c1-200
  tail -3 gpurun_out/bench_n${N}${mode}.err | cut -c1-300
done

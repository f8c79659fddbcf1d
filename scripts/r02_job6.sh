#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_job6.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1_job6.json 2> gpurun_out/bench_n1_job6.err
tail -c 6000 gpurun_out/bench_n1_job6.json; tail -5 gpurun_out/bench_n1_job6.err

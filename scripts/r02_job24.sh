#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_job24.log
python scripts/skew_probe.py 24 2>&1 | tee gpurun_out/skew_24.log

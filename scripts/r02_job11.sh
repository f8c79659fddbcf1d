#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "reduce_scatter or sharding or golden or smoke" 2>&1 | tail -5
bash scripts/r02_job7.sh 8

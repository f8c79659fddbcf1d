#!/bin/bash
mkdir -p gpurun_out
{
python scripts/sweep.py 0:26 0:24 0:23 0:22 1:24 2>&1 | cut -c1-420
echo "== skew"; python scripts/skew_probe.py 24 2>&1 | cut -c1-300
} > gpurun_out/scatter_unrolled.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3

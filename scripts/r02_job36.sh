#!/bin/bash
mkdir -p gpurun_out
{
echo "== default (whole-bin blocks)"; python scripts/sweep.py 0:26 0:24 0:23 1:24 2>&1 | cut -c1-420
echo "== B200MSM_SCATTER=tiled"; B200MSM_SCATTER=tiled python scripts/sweep.py 0:26 0:23 2>&1 | cut -c1-420
echo "== skew"; python scripts/skew_probe.py 24 2>&1 | cut -c1-300
} > gpurun_out/whole_scatter.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3

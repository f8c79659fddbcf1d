#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_bin_scatter_tiled|k_digits' -s 6 -c 3 -f -o gpurun_out/r02_scatter_kernels python scripts/sweep.py 0:26 > gpurun_out/ncu_scatter.log 2>&1
tail -3 gpurun_out/ncu_scatter.log | cut -c1-300
ls -la gpurun_out/*.ncu-rep

#!/bin/bash
# ncu: k_accumulate (default build = lazy-norm) at 2^22, full set with source; plus the launch list of one 2^26 MSM
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_accumulate -s 4 -c 1 -f -o gpurun_out/r02_acc python scripts/sweep.py 0:22 > gpurun_out/ncu_acc.log 2>&1
tail -2 gpurun_out/ncu_acc.log | cut -c1-300
ncu --set full --clock-control none --import-source on -k regex:k_reduce_slices -s 6 -c 1 -f -o gpurun_out/r02_red0 python scripts/sweep.py 0:24 > gpurun_out/ncu_red.log 2>&1
tail -1 gpurun_out/ncu_red.log | cut -c1-300
ls -la gpurun_out/*.ncu-rep

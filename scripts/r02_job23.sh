#!/bin/bash
mkdir -p gpurun_out
bash scripts/r02_variants.sh dflt tma e160
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active"
for v in dflt tma e160; do
  echo "== ncu $v (2^24, k_accumulate)"
  B200MSM_LIB=$PWD/variants/lib_$v.so ncu --metrics $M --clock-control none -k regex:k_accumulate -s 4 -c 1 --csv python scripts/sweep.py 0:24 2>/dev/null | grep -E "^\"[0-9]" | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' | tr -d '"'
done
# the shipped build: full-set capture of k_accumulate at the headline size, and the launch list of one 2^26 MSM
ncu --set full --clock-control none --import-source on -k regex:k_accumulate -s 4 -c 1 -f -o gpurun_out/r02_acc26 python scripts/sweep.py 0:26 > gpurun_out/ncu_acc26.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 30 --csv --log-file gpurun_out/r02_launches26.csv python scripts/sweep.py 0:26 > /dev/null 2>&1
ls -la gpurun_out/r02_acc26.ncu-rep gpurun_out/r02_launches26.csv

#!/bin/bash
# GPU job 1 of round 2: (a) the reference's fastest entries re-timed on this B200, (b) accumulate-kernel variants at 2^24.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/job1_gpu.txt 2>&1
python scripts/retime_reference_entry.py 20 4 100 3 > gpurun_out/retime_20.log 2>&1
cat gpurun_out/retime_20.log | grep -v '"stage"'
ok=b200msm
grep -q '"impl": "combined".*best_ms_total' gpurun_out/retime_20.log && ok=combined,$ok
grep -q '"impl": "yrrid".*best_ms_total' gpurun_out/retime_20.log && ok=yrrid,$ok
echo "2^26 run for: $ok"
python scripts/retime_reference_entry.py 26 4 400 3 $ok > gpurun_out/retime_26.log 2>&1
cat gpurun_out/retime_26.log | grep -v '"stage"'
for v in base lazy1 lazy2 inline inline_lazy inline_b2 tpb64 tpb96; do
  if [ -f variants/lib_$v.so ]; then
    echo "== variant $v" | tee -a gpurun_out/variants_24.log
    B200MSM_LIB=$PWD/variants/lib_$v.so timeout 300 python scripts/sweep.py 0:24 2>&1 | tail -2 | tee -a gpurun_out/variants_24.log
  fi
done

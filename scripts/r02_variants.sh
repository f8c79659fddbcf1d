#!/bin/bash
# accumulate-kernel variants (scripts/build_variants.py) at 2^lg: correctness against the closed form + stage times
mkdir -p gpurun_out
LG=${LG:-24}
for v in "$@"; do
  if [ -f variants/lib_$v.so ]; then
    echo "== variant $v" | tee -a gpurun_out/variants.log
    B200MSM_LIB=$PWD/variants/lib_$v.so timeout 300 python scripts/sweep.py 0:$LG 2>&1 | tail -1 | python -c "
import sys,json
for l in sys.stdin:
    try:
        d=json.loads(l); print({k:d[k] for k in ('ms_per_msm','bit_exact_vs_oracle')}, d['stages_ms'])
    except Exception: print(l[:300])
" | tee -a gpurun_out/variants.log
  fi
done

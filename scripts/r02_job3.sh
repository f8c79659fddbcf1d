#!/bin/bash
# GPU job 3 of round 2: reference entry with the shared-window rebase; then the GPU parity suite.
mkdir -p gpurun_out
echo "== diag libcombined_msm_diag.so (smem-base 1024)" | tee gpurun_out/ref_diag2.log
timeout 90 python scripts/ref_entry_diag.py oracle/_ref/libcombined_msm_diag.so 20 2>&1 | tee -a gpurun_out/ref_diag2.log
if grep -q "diag rc 0" gpurun_out/ref_diag2.log; then
  python scripts/retime_reference_entry.py 20 4 100 3 2>&1 | grep -v '"stage"' | tee gpurun_out/retime2_20.log
  ok=b200msm
  grep -q '"impl": "combined".*best_ms_total' gpurun_out/retime2_20.log && ok=combined,$ok
  grep -q '"impl": "yrrid".*best_ms_total' gpurun_out/retime2_20.log && ok=yrrid,$ok
  python scripts/retime_reference_entry.py 26 4 400 3 $ok 2>&1 | grep -v '"stage"' | tee gpurun_out/retime2_26.log
fi
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_job3.log

#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  echo "compute-sanitizer --tool $tool python scripts/sanitize_probe.py"
  timeout 900 compute-sanitizer --tool $tool python scripts/sanitize_probe.py 2>&1 | grep -v "^=========     \|Host Frame" | tail -16
done | tee gpurun_out/sanitizer_r02.log

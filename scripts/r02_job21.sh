#!/bin/bash
show() { python -c "
import sys,json
for l in sys.stdin:
    try:
        d=json.loads(l); print(d['ms_per_msm'], d['bit_exact_vs_oracle'], d['stages_ms']['scan_scatter'])
    except Exception: print(l[:200])"; }
for k in 32 64 256 512 1024; do
echo "== 2^26 tile ${k}K"; B200MSM_SCATTER_TILE_K=$k python scripts/sweep.py 0:26 2>&1 | tail -1 | show
done

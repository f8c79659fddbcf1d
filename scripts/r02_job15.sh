#!/bin/bash
for mb in 512 128; do for lg in 21 22 23; do
echo "== BIN_MIN_MB=$mb 2^$lg"; B200MSM_BIN_MIN_MB=$mb python scripts/sweep.py 0:$lg 2>&1 | tail -1 | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['ms_per_msm'], d['bit_exact_vs_oracle'], d['plan']['window_bits'], d['stages_ms'])"
done; done

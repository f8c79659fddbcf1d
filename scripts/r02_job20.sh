#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q -k "two_level or three_tables or closed_form_2pow20 or edge_cases or golden or chunked" 2>&1 | tail -4
show() { python -c "
import sys,json
for l in sys.stdin:
    try:
        d=json.loads(l); print(d['ms_per_msm'], d['bit_exact_vs_oracle'], d['stages_ms'])
    except Exception: print(l[:200])"; }
for lg in 24 26; do
echo "== 2^$lg atomic";  B200MSM_SCATTER=atomic python scripts/sweep.py 0:$lg 2>&1 | tail -1 | show
echo "== 2^$lg tiled 24MB bins"; python scripts/sweep.py 0:$lg 2>&1 | tail -1 | show
echo "== 2^$lg tiled 12MB bins"; B200MSM_BIN_KB=12288 python scripts/sweep.py 0:$lg 2>&1 | tail -1 | show
done

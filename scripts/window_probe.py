"""Stage times for forced window widths (plan sanity): python scripts/window_probe.py LOG_N C [C ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import __graft_entry__ as g
M = g.load_package()
lg = int(sys.argv[1]); n = 1 << lg
for c in [int(a) for a in sys.argv[2:]]:
    e = M.Engine(0, window_bits=c)
    e.set_bases_generated(n, seed=77, mode=0, distinct=1 << 15)
    sc = torch.empty(n * 32, dtype=torch.uint8, device="cuda"); out = torch.zeros(144, dtype=torch.uint8, device="cuda")
    e.generate_scalars(sc.data_ptr(), n, 5); e.sync()
    for _ in range(3):
        e.msm_device(sc.data_ptr(), n, out.data_ptr())
    e.sync()
    i = e.info()
    print(c, {k: i[k] for k in ("windows", "tables", "buckets", "lmax")}, {k: round(v, 2) for k, v in e.stage_ms().items()}, flush=True)
    e.close()

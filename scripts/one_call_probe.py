"""Times the one-call form (VariableBaseMSM.multi_scalar_mul: bases AND scalars from host memory on every call) over sizes
and first-call chunk counts, each result checked against the closed form.  Usage: one_call_probe.py [max_log_n]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import __graft_entry__ as g
from oracle import oracle as O
import known_answer as KA
M = g.load_package()
max_lg = int(sys.argv[1]) if len(sys.argv) > 1 else 24
n = 1 << max_lg
SEED = 77
eng = M.Engine(0, tables=1)
eng.set_bases_generated(n, seed=SEED, mode=M.BASES_HARNESS, distinct=1 << 15)
hb = eng.get_bases(0, n)
sc = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
eng.generate_scalars(sc.data_ptr(), n, 5)
eng.sync()
hs_pinned = torch.empty(n * 32, dtype=torch.uint8, pin_memory=True)
hs_pinned.copy_(sc)
torch.cuda.synchronize()
del sc
eng.close()
hs = hs_pinned.numpy()
hs_pageable = hs.copy()
logs = KA.base_logs(n, SEED, 0, 1 << 15)
r_mod = O.constants(0)["r"]
V = M.VariableBaseMSM
sizes = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else (16, 20, 22, 24, 26)
variants = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else (0, 1, 2, 8)
for lg in [x for x in sizes if x <= max_lg]:
    m = 1 << lg
    exp = KA.expected_affine(O, 0, KA.dot_mod_r(hs[:m * 32].reshape(m, 32), logs[:m], r_mod))
    for chunks in ((0,) if lg < 22 else variants):
        for pinned in ((True, False) if chunks == 0 else (True,)):
            if chunks: os.environ["B200MSM_HOST_CHUNKS"] = str(chunks)
            else: os.environ.pop("B200MSM_HOST_CHUNKS", None)
            V.release()
            s = (hs if pinned else hs_pageable)[:m * 32]
            out = V.multi_scalar_mul(hb[:m], s, stride=96)
            ok = bool(np.array_equal(O.jac_to_affine(out)[0], exp))
            ts = []
            for _ in range(3):
                t0 = time.perf_counter(); V.multi_scalar_mul(hb[:m], s, stride=96); ts.append((time.perf_counter() - t0) * 1e3)
            print(json.dumps({"log_n": lg, "chunks": chunks or "auto", "scalars": "pinned" if pinned else "pageable", "bases": "pageable", "ms_min": round(min(ts), 2), "ms_all": [round(t, 1) for t in ts], "bit_exact": ok}), flush=True)
V.release()

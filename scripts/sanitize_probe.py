"""Small MSMs through every pipeline variant, for compute-sanitizer (memcheck / racecheck) runs on the GPU box."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import __graft_entry__ as g
from oracle import oracle as O
M = g.load_package()
n = 1 << 11
for curve in (0, 1):
    pts = np.tile(O.gen_points(64, seed=3, curve=curve), (n // 64, 1))
    sc = O.gen_scalars(n * 2, seed=4, curve=curve)
    for wb, chunks, binmb in ((0, 0, None), (16, 3, "0")):
        if binmb is not None: os.environ["B200MSM_BIN_MIN_MB"] = binmb; os.environ["B200MSM_BIN_KB"] = "1"     # 1-KB bins: really two-level at this size
        else: os.environ.pop("B200MSM_BIN_MIN_MB", None); os.environ.pop("B200MSM_BIN_KB", None)
        os.environ["B200MSM_HOST_CHUNKS"] = str(chunks)
        e = M.Engine(curve, window_bits=wb)
        e.set_bases(pts)
        out = e.msm_host(sc, 2)
        ok = all(np.array_equal(O.jac_to_affine(out[b], curve), O.jac_to_affine(O.ark_msm(pts, sc[b*n:(b+1)*n], curve), curve)) for b in range(2))
        print("curve", curve, "window", wb, "chunks", chunks, "two-level asked", binmb is not None, "ok", ok, e.info(), flush=True)
        e.close()

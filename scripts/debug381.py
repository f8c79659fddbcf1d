import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import __graft_entry__ as g
from oracle import oracle as O
M = g.load_package()
def run(curve, pts, sc, c, T):
    eng = M.Engine(curve, window_bits=c, tables=T)
    eng.set_bases(pts)
    n = len(sc)
    s = torch.from_numpy(np.ascontiguousarray(sc).reshape(-1)).cuda()
    out = torch.zeros(144, dtype=torch.uint8, device="cuda")
    eng.msm_device(s.data_ptr(), n, out.data_ptr()); eng.sync()
    info = eng.info(); eng.close()
    got = O.jac_to_affine(out.cpu().numpy(), curve)[0]
    exp = O.jac_to_affine(O.ark_msm(pts, sc, curve), curve)[0]
    return np.array_equal(got, exp), info
curve = 1
n = 1 << 12
for inf in (0, 1):
  for neg in (0, 1):
    pts = np.tile(O.gen_points(128, seed=31 + curve, curve=curve), (n // 128, 1))
    sc = O.gen_scalars(n, seed=32, curve=curve)
    if inf:
        pts[5] = 0; pts[5, 96] = 1
    if neg:
        pts[9] = O.aff_neg(pts[8:9], curve)[0]; sc[9] = sc[8]
    for c, T in ((0, 0), (0, 1), (13, 1), (13,0), (8, 0), (16, 0)):
        ok, info = run(curve, pts, sc, c, T)
        print(f"inf={inf} neg={neg} c={c} T={T} -> ok={ok} plan c={info['window_bits']} W={info['windows']} T={info['tables']} F={info['folded_windows']}", flush=True)
# minimal: two points P, -P with equal scalar + one more
pts = O.gen_points(4, seed=3, curve=curve); sc = O.gen_scalars(4, seed=4, curve=curve)
pts[1] = O.aff_neg(pts[0:1], curve)[0]; sc[1] = sc[0]
for c, T in ((0, 0), (0, 1), (4, 1), (8, 1)):
    ok, info = run(curve, pts, sc, c, T)
    print(f"minimal neg c={c} T={T} -> ok={ok} {info}", flush=True)

import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import __graft_entry__ as g
from oracle import oracle as O
M = g.load_package()
p = O.constants(0)["p"]
n = 1 << 11
pts = np.tile(O.gen_points(256, seed=71), (n // 256, 1))
sc = O.gen_scalars(n, seed=72)
e = M.Engine(0); e.set_bases(pts)
s = torch.from_numpy(sc.reshape(-1).copy()).cuda()
part = torch.zeros(224, dtype=torch.uint8, device="cuda")
e.msm_partial_device(s.data_ptr(), n, part.data_ptr()); e.sync()
raw = part.cpu().numpy().view(np.uint32)
vals = [sum(int(raw[14*k+i]) << (28*i) for i in range(14)) for k in range(4)]
X, Y, ZZ, ZZZ = vals
print("limbs X", [hex(int(v)) for v in raw[:14]])
x = X * pow(ZZ, -1, p) % p; y = Y * pow(ZZZ, -1, p) % p
exp = O.jac_to_affine(O.ark_msm(pts, sc))[0]
R = 1 << 384
ex = int.from_bytes(bytes(exp[:48]), "little") * pow(R, -1, p) % p
ey = int.from_bytes(bytes(exp[48:96]), "little") * pow(R, -1, p) % p
print("partial x ok", x == ex, "y ok", y == ey, "zz3==zzz2", pow(ZZ,3,p) == pow(ZZZ,2,p))
out = torch.zeros(144, dtype=torch.uint8, device="cuda")
e.combine_device(part.data_ptr(), 1, out.data_ptr()); e.sync()
o = out.cpu().numpy()
jx, jy, jz = [int.from_bytes(bytes(o[48*k:48*k+48]), "little") * pow(R, -1, p) % p for k in range(3)]
print("combine jac: x/z^2 ok", jx * pow(jz, -2, p) % p == ex, "y/z^3 ok", jy * pow(jz, -3, p) % p == ey)
print("jz == ZZ/R'?", jz, ZZ * pow(1 << 392, -1, p) % p)
print("jx vs X*ZZ", jx == (X * ZZ * pow(1 << 392, -2, p)) % p)

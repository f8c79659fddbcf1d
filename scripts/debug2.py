import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import __graft_entry__ as g
from oracle import oracle as O
M = g.load_package()
n, batches = 1 << 10, 4
pts = np.tile(O.gen_points(64, seed=61), (n // 64, 1))
sc = O.gen_scalars(n * batches, seed=62)
y = M.YrridContext(max_points=1 << 12, max_batches=16)
y.preprocess_points(pts)
out = y.run(sc)
for b in range(batches):
    print("yrrid batch", b, np.array_equal(O.jac_to_affine(out[b]), O.jac_to_affine(O.ark_msm(pts, sc[b * n:(b + 1) * n]))))
y.free()
ctx = M.multi_scalar_mult_init(pts)
out = M.multi_scalar_mult(ctx, pts, sc)
for b in range(batches):
    print("famA batch", b, np.array_equal(O.jac_to_affine(out[b]), O.jac_to_affine(O.ark_msm(pts, sc[b * n:(b + 1) * n]))))
e = M.Engine(0); e.set_bases(pts)
for b in range(batches):
    o = e.msm_host(sc[b*n:(b+1)*n], 1)
    print("engine single batch", b, np.array_equal(O.jac_to_affine(o[0]), O.jac_to_affine(O.ark_msm(pts, sc[b * n:(b + 1) * n]))))
o = e.msm_host(sc, 4)
for b in range(batches):
    print("engine 4 batches", b, np.array_equal(O.jac_to_affine(o[b]), O.jac_to_affine(O.ark_msm(pts, sc[b * n:(b + 1) * n]))))
# sharding
n = 1 << 12
pts = np.tile(O.gen_points(256, seed=71), (n // 256, 1))
sc = O.gen_scalars(n, seed=72)
parts = torch.zeros(2 * M.PARTIAL_BYTES, dtype=torch.uint8, device="cuda")
engs = []
for gI in range(2):
    e = M.Engine(0)
    e.set_bases(pts[gI * n // 2:(gI + 1) * n // 2])
    s = torch.from_numpy(sc[gI * n // 2:(gI + 1) * n // 2].reshape(-1).copy()).cuda()
    e.msm_partial_device(s.data_ptr(), n // 2, parts.data_ptr() + gI * M.PARTIAL_BYTES)
    e.sync()
    o2 = torch.zeros(144, dtype=torch.uint8, device="cuda")
    e.msm_device(s.data_ptr(), n // 2, o2.data_ptr()); e.sync()
    print("shard", gI, "full ok", np.array_equal(O.jac_to_affine(o2.cpu().numpy()), O.jac_to_affine(O.ark_msm(pts[gI * n // 2:(gI + 1) * n // 2], sc[gI * n // 2:(gI + 1) * n // 2]))))
    print("partial nonzero bytes", int((parts.cpu().numpy()[gI*224:(gI+1)*224] != 0).sum()))
    engs.append(e)
out = torch.zeros(144, dtype=torch.uint8, device="cuda")
engs[0].combine_device(parts.data_ptr(), 2, out.data_ptr()); engs[0].sync()
print("combined nonzero", int((out.cpu().numpy() != 0).sum()), np.array_equal(O.jac_to_affine(out.cpu().numpy()), O.jac_to_affine(O.ark_msm(pts, sc))))
engs[0].combine_device(parts.data_ptr(), 1, out.data_ptr()); engs[0].sync()
print("combine count=1 == shard0", np.array_equal(O.jac_to_affine(out.cpu().numpy()), O.jac_to_affine(O.ark_msm(pts[:n//2], sc[:n//2]))))

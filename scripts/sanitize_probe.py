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

# round 2: the multi-GPU bucket reduce-scatter (k_peer_merge with TMA bulk copies, ready / done counters) between two contexts of
# this process, and the XYZZ arithmetic of BLS12-377
import torch
for k in ("B200MSM_BIN_MIN_MB", "B200MSM_BIN_KB", "B200MSM_HOST_CHUNKS"):
    os.environ.pop(k, None)
world, n_local = 2, 1 << 11
pts = np.tile(O.gen_points(64, seed=5), (world * n_local // 64, 1))
engines = []
for r in range(world):
    e = M.Engine(0, window_bits=10)
    e.set_bases(pts[r * n_local:(r + 1) * n_local])
    engines.append(e)
handles = np.concatenate([e.peer_export(1) for e in engines])
for r, e in enumerate(engines):
    e.peer_attach(r, world, handles)
parts = torch.zeros(world * 224, dtype=torch.uint8, device="cuda")
out = torch.zeros(144, dtype=torch.uint8, device="cuda")
for epoch in range(2):
    sc = O.gen_scalars(world * n_local, seed=40 + epoch)
    dev = torch.from_numpy(sc.reshape(-1).copy()).cuda()
    for r, e in enumerate(engines):
        e.peer_accumulate_device(dev[r * n_local * 32:].data_ptr(), n_local)
    for e in engines:
        e.sync()
    for r, e in enumerate(engines):
        e.peer_reduce(parts[r * 224:].data_ptr())
    for e in engines:
        e.sync()
    engines[0].combine_device(parts.data_ptr(), world, out.data_ptr())
    engines[0].sync()
    print("bucket reduce-scatter, 2 contexts, epoch", epoch, "ok", bool(np.array_equal(O.jac_to_affine(out.cpu().numpy()), O.jac_to_affine(O.ark_msm(pts, sc)))), flush=True)
for e in engines:
    e.peer_detach(); e.close()
e = M.Engine(0)
e.set_bucket_arithmetic(True)
e.set_bases(pts)
sc = O.gen_scalars(len(pts), seed=50)
print("BLS12-377 XYZZ arithmetic ok", bool(np.array_equal(O.jac_to_affine(e.msm_host(sc)), O.jac_to_affine(O.ark_msm(pts, sc)))), flush=True)
e.close()

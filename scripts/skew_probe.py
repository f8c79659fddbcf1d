"""Behaviour on skewed scalar sets: time and check a few degenerate distributions at 2^24."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import __graft_entry__ as g
from oracle import oracle as O
import known_answer as KA
M = g.load_package()
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 24
n = 1 << lg
r = O.constants(0)["r"]
e = M.Engine(0)
e.set_bases_generated(n, seed=77, mode=0, distinct=1 << 15)
logs = KA.base_logs(n, 77, 0, 1 << 15)
rng = np.random.default_rng(1)
def run(name, sc_np):
    sc = torch.from_numpy(sc_np.reshape(-1)).cuda()
    out = torch.zeros(144, dtype=torch.uint8, device="cuda")
    e.msm_device(sc.data_ptr(), n, out.data_ptr()); e.sync()
    t0 = time.time(); e.msm_device(sc.data_ptr(), n, out.data_ptr()); e.sync(); dt = time.time() - t0
    t = KA.dot_mod_r(sc_np, logs, r)
    ok = np.array_equal(O.jac_to_affine(out.cpu().numpy())[0], KA.expected_affine(O, 0, t))
    print(f"{name:28s} ok={ok} {dt*1e3:9.2f} ms  stages={ {k: round(v, 2) for k, v in e.stage_ms().items()} }", flush=True)
one = np.frombuffer(int(0x0123456789abcdef0123456789abcdef0123456789abcdef0123456789abcdef % r).to_bytes(32, "little"), np.uint8)
run("all equal", np.tile(one, (n, 1)))
small = np.zeros((n, 32), np.uint8); small[:, 0] = rng.integers(0, 256, n, dtype=np.uint8)
run("8-bit scalars", small)
two = np.zeros((n, 32), np.uint8); two[:, :4] = rng.integers(0, 256, (n, 4), dtype=np.uint8)
run("32-bit scalars", two)
sparse = np.zeros((n, 32), np.uint8); idx = rng.choice(n, n // 64, replace=False); sparse[idx] = rng.integers(0, 256, (len(idx), 32), dtype=np.uint8); sparse[:, 31] &= 0x0f
run("1/64 non-zero", sparse)
zeros = np.zeros((n, 32), np.uint8)
run("all zero", zeros)

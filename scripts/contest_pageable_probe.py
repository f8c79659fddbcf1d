"""The contest's timed call (one multi_scalar_mult with 4 batches of 2^26 scalars) with the scalars in PAGEABLE host memory,
as the Rust harness passes them (a Vec), with and without the staged upload (B200MSM_NO_STAGING=1 = plain cudaMemcpyAsync)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import __graft_entry__ as g
M = g.load_package()
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 26
n, batches = 1 << lg, 4
eng = M.Engine(0)
eng.set_bases_generated(n, seed=77, mode=M.BASES_HARNESS, distinct=1 << 15)
dev = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
pageable = np.empty(batches * n * 32, np.uint8)
for b in range(batches):
    eng.generate_scalars(dev.data_ptr(), n, 200 + b)
    eng.sync()
    pageable[b * n * 32:(b + 1) * n * 32] = dev.cpu().numpy()
pinned = torch.empty(batches * n * 32, dtype=torch.uint8, pin_memory=True)
pinned.numpy()[:] = pageable
ref = None
for name, buf, env in (("pinned", pinned.numpy(), None), ("pageable, staged through pinned slots", pageable, None), ("pageable, plain cudaMemcpyAsync", pageable, "1")):
    if env: os.environ["B200MSM_NO_STAGING"] = env
    else: os.environ.pop("B200MSM_NO_STAGING", None)
    out = eng.msm_host(buf, batches)
    ts = []
    for _ in range(2):
        t0 = time.perf_counter(); out = eng.msm_host(buf, batches); ts.append((time.perf_counter() - t0) * 1e3)
    if ref is None: ref = out.copy()
    print(json.dumps({"log_n": lg, "batches": batches, "scalars": name, "ms_total_min": round(min(ts), 1), "ms_all": [round(t, 1) for t in ts], "same_result": bool(np.array_equal(out, ref))}), flush=True)
eng.close()

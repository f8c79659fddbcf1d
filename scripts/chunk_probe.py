"""End-to-end time of multi_scalar_mult (one batch, pinned host scalars) for chunk counts / size ratios of the first batch.
  python scripts/chunk_probe.py [log_n=26] chunks:ratio ..."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import __graft_entry__ as g
M = g.load_package()
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 26
n = 1 << lg
for cfg in sys.argv[2:] or ["4:2", "3:4", "3:5", "4:3", "2:5"]:
    ch, q = cfg.split(":")
    os.environ["B200MSM_HOST_CHUNKS"], os.environ["B200MSM_CHUNK_RATIO"] = ch, q
    e = M.Engine(0)
    e.set_bases_generated(n, seed=77, mode=0, distinct=1 << 15)
    sc = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
    e.generate_scalars(sc.data_ptr(), n, 5)
    e.sync()
    host = torch.empty(n * 32, dtype=torch.uint8, pin_memory=True)
    host.copy_(sc); torch.cuda.synchronize()
    del sc
    ref = None
    ts = []
    for it in range(6):
        t0 = time.perf_counter(); r = e.msm_host(host.data_ptr(), 1); ts.append((time.perf_counter() - t0) * 1e3)
        ref = r if ref is None else ref
        assert np.array_equal(r, ref)
    print(f"2^{lg} chunks {ch} ratio {q}: e2e ms {[round(t, 1) for t in ts]}  best {min(ts):.1f}", flush=True)
    e.close(); del host; torch.cuda.empty_cache()

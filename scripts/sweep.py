"""BASELINE configs 2 and 4: size sweep 2^20..2^28 on BLS12-377 (window plan auto-tuned per size) and BLS12-381 at 2^24,
each checked against the closed form (tests/known_answer.py) and timed with CUDA events (3 warm-up + 5 timed MSMs).
Usage (GPU box): python scripts/sweep.py [curve:log_n ...]   default: 0:20 0:22 0:24 0:26 0:28 1:24"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import __graft_entry__ as g
from oracle import oracle as O
import known_answer as KA
M = g.load_package()
cases = [tuple(int(v) for v in a.split(":")) for a in sys.argv[1:]] or [(0, 20), (0, 22), (0, 24), (0, 26), (0, 28), (1, 24)]
for curve, lg in cases:
    n = 1 << lg
    try:
        e = M.Engine(curve)
        e.set_bases_generated(n, seed=77, mode=0, distinct=1 << 15)
        s = torch.cuda.Stream()
        e.set_stream(s.cuda_stream)
        sc = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
        out = torch.zeros(144, dtype=torch.uint8, device="cuda")
        e.generate_scalars(sc.data_ptr(), n, 5)
        e.sync()
        for _ in range(3):
            e.msm_device(sc.data_ptr(), n, out.data_ptr())
        e.sync(); e.reset_stage_stats()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(5):
            e.msm_device(sc.data_ptr(), n, out.data_ptr())
        e1.record(s); e.sync()
        ms = e0.elapsed_time(e1) / 5
        st, k = e.stage_ms_sum()
        got = O.jac_to_affine(out.cpu().numpy(), curve)[0]
        ok = None
        if lg <= 26:      # the closed-form check needs the scalars on the host
            t = KA.dot_mod_r(sc.cpu().numpy().reshape(n, 32), KA.base_logs(n, 77, 0, 1 << 15), O.constants(curve)["r"])
            ok = bool(np.array_equal(got, KA.expected_affine(O, curve, t)))
        info = e.info()
        print(json.dumps({"curve": ["BLS12-377", "BLS12-381"][curve], "log_n": lg, "ms_per_msm": round(ms, 3), "points_per_s": round(n / ms * 1e3),
                          "bit_exact_vs_oracle": ok, "plan": {k2: info[k2] for k2 in ("window_bits", "windows", "tables", "folded_windows", "buckets")},
                          "stages_ms": {a: round(v / max(k, 1), 3) for a, v in st.items()}, "hbm_GBps_algorithmic": round(n * 128 / ms / 1e6, 1)}), flush=True)
        e.close(); del sc, out
        torch.cuda.empty_cache()
    except Exception as ex:
        print(json.dumps({"curve": curve, "log_n": lg, "error": str(ex)[:200]}), flush=True)

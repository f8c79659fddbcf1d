"""Window-plan calibration: ms per MSM for forced window widths at shard sizes (1 GPU), against choose_plan's own choice.
  python scripts/plan_probe.py log_n:c ...      (c = 0: the planner's choice)"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import __graft_entry__ as g
M = g.load_package()
for a in sys.argv[1:]:
    lg, c = (int(v) for v in a.split(":"))
    n = 1 << lg
    e = M.Engine(0, window_bits=c)
    e.set_bases_generated(n, seed=77, mode=0, distinct=1 << 15)
    sc = torch.empty(n * 32, dtype=torch.uint8, device="cuda"); out = torch.zeros(144, dtype=torch.uint8, device="cuda")
    e.generate_scalars(sc.data_ptr(), n, 5); e.sync()
    ref = None
    for _ in range(3): e.msm_device(sc.data_ptr(), n, out.data_ptr())
    e.sync(); e.reset_stage_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st = torch.cuda.current_stream()
    torch.cuda.synchronize(); e0.record()
    for _ in range(5): e.msm_device(sc.data_ptr(), n, out.data_ptr())
    e.sync(); e1.record(); torch.cuda.synchronize()
    stg, k = e.stage_ms_sum(); info = e.info()
    print(json.dumps({"log_n": lg, "forced_c": c, "c": info["window_bits"], "windows": info["windows"], "buckets": info["buckets"], "ms": round(e0.elapsed_time(e1) / 5, 3),
                      "stages": {a: round(v / k, 2) for a, v in stg.items()}, "result": out.cpu().numpy()[:8].tolist()}), flush=True)
    e.close(); del sc; torch.cuda.empty_cache()

"""First-contact probe on the GPU box: correctness of generated-bases MSMs vs the closed form, stage timings."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import __graft_entry__ as g
from oracle import oracle as O
import known_answer as KA
M = g.load_package()
sizes = [int(a) for a in sys.argv[1:]] or [12, 16, 20]
for lg in sizes:
    n = 1 << lg
    for mode in (0, 1):
        t0 = time.time()
        e = M.Engine()
        e.set_bases_generated(n, seed=77, mode=mode, distinct=1 << 15)
        t_init = time.time() - t0
        sc = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
        out = torch.zeros(144, dtype=torch.uint8, device="cuda")
        e.generate_scalars(sc.data_ptr(), n, 5)
        e.sync()
        for it in range(3):
            t1 = time.time(); e.msm_device(sc.data_ptr(), n, out.data_ptr()); e.sync(); dt = time.time() - t1
        st = e.stage_ms()
        got = O.jac_to_affine(out.cpu().numpy())[0]
        logs = KA.base_logs(n, 77, mode, 1 << 15)
        t = KA.dot_mod_r(sc.cpu().numpy().reshape(n, 32), logs, O.constants(0)["r"])
        exp = KA.expected_affine(O, 0, t)
        ok = np.array_equal(got, exp)
        print(f"2^{lg} mode={mode} ok={ok} init={t_init:.2f}s msm={dt*1e3:.2f}ms info={e.info()} stages={ {k: round(v, 3) for k, v in st.items()} }", flush=True)
        e.close()
        del sc, out
        torch.cuda.empty_cache()

"""End-to-end time of ShardedMsm.msm (pinned host scalars per rank) for chunk counts / ratios of the upload, N ranks.
  torchrun --nproc-per-node N scripts/sharded_e2e_probe.py chunks:ratio ..."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import importlib, numpy as np, torch, torch.distributed as dist
import __graft_entry__ as g
M = g.load_package()
SH = importlib.import_module(g.PKG_NAME + ".sharded")
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
if "NCCL_DEBUG_FILE" not in os.environ:
    os.environ["NCCL_DEBUG"] = "WARN"; os.environ["NCCL_DEBUG_FILE"] = "/dev/stderr"
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = (1 << 26) // world
eng = M.Engine(0, device=local)
eng.set_bases_generated(n, seed=77, mode=0, distinct=1 << 15)
sh = SH.ShardedMsm(SH.CudaShardBackend(eng))
scal = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
eng.generate_scalars(scal.data_ptr(), n, 1234 + rank); eng.sync()
host = torch.empty(n * 32, dtype=torch.uint8, pin_memory=True); host.copy_(scal); torch.cuda.synchronize()
ref = None
for cfg in sys.argv[1:]:
    ch, q = (int(v) for v in cfg.split(":"))
    eng.set_host_chunks(ch, q)
    sh.attach_peers(chunks=ch)
    for _ in range(2):
        r = sh.msm(host)
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(6):
        r = sh.msm(host)
    dist.barrier(); torch.cuda.synchronize()
    ms = torch.tensor([(time.perf_counter() - t0) / 6 * 1e3], dtype=torch.float64, device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ref = r.clone() if ref is None else ref
    if rank == 0:
        print(f"N={world} chunks {ch} ratio {q}: e2e {ms.item():.2f} ms  same result {bool(torch.equal(r, ref))}", flush=True)
sh.detach_peers()
dist.destroy_process_group()

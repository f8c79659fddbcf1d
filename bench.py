#!/usr/bin/env python3
"""bench.py -- BLS12-377 G1 MSM at 2^26 points on 1/2/4/8 B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--log-n 26] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One step = one MSM (one batch of the harness): n = 2^log_n bases resident on the GPU(s) (init is untimed, as in the
contest: multi_scalar_mult_init), scalars -> one G1Projective.  With N ranks the point set is sharded evenly
(2^log_n / N per GPU, "strong" scaling), each rank runs the whole single-GPU pipeline on its slice, and the per-rank
partial sums are exchanged by one NCCL all-gather (2022-entries_b200/sharded.py).

value      points/s, scalars already in HBM when the timed region starts (CUDA events, max over ranks)
e2e        the same through the harness call multi_scalar_mult (C ABI mult_pippenger_inf at N=1) with the scalars in
           pinned HOST memory: H2D of the batch and D2H of the result are inside the timed region
roofline   the accumulate kernel (dominant): algorithmic bytes = n * 128 B (96-B base + 32-B scalar) over its
           measured duration vs the measured HBM copy peak.  The kernel is integer-pipe bound by construction
           (DESIGN.md), so the int_pipe block next to it is the number that says how good the kernel is.
cpu_baseline  the CPU oracle (C restatement of arkworks VariableBaseMSM) on this box's host cores, bounded sample
--impl reference   times that CPU path alone (the reference's CPU implementation is Rust/arkworks: no toolchain here,
           so the port stands in; kind = "port")
Correctness of every timed configuration is asserted before timing: the bases are k_i * G with known k_i, so the
result must equal (sum s_i k_i mod r) * G, computed by the CPU oracle (tests/known_answer.py).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "BLS12-377 G1 MSM 2^{lg} points/s"
WORKLOAD = ("BLS12-377 G1 MSM 2^{lg}, one batch, bases resident (harness generator shape: 2^15 distinct points replicated), "
            "uniform 253-bit scalars")
SEED_BASES, SEED_SCALARS, DISTINCT = 77, 1234, 1 << 15
TRAFFIC_FILE = "r02_traffic.json"      # ncu --set full capture behind roofline.traffic
REF_SAMPLE_LOG_N = 22          # the CPU arm's sample, the same on every host


def measured_peak_gbs():
    """HBM copy bandwidth the driver measured on this pool (MEASURED_PEAKS.json), else the recipe's fallback."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")

    def find(d):
        if isinstance(d, dict):
            for k, v in d.items():           # e.g. {"hbm_gbs": 6451.8, ...}; tolerate nesting and unit suffixes
                if "hbm" in str(k).lower() and isinstance(v, (int, float)) and v > 0:
                    return float(v) * (1000.0 if v < 50 else 1.0)      # a value in TB/s
                r = find(v)
                if r:
                    return r
        return None
    try:
        with open(path) as f:
            v = find(json.load(f))
        if v:
            return v, "measured (MEASURED_PEAKS.json)"
    except Exception:
        pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU during the timed region (pynvml; nvidia-smi's counters)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                     "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                     "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                     "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
            while not self.stop_flag:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
                time.sleep(0.05)
        except Exception as ex:      # no NVML: report it rather than invent clocks
            self.reasons.add(f"nvml_unavailable:{type(ex).__name__}")

    def result(self):
        self.stop_flag = True
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def cpu_oracle_rate(O, log_sample, threads=None):
    """points/s of the CPU oracle (arkworks algorithm) on a harness-shaped sample of 2^log_sample points."""
    import numpy as np
    n = 1 << log_sample
    base = O.gen_points(1 << 10, seed=3)
    P = np.tile(base, (n >> 10, 1))
    S = O.gen_scalars(n, seed=4)
    t0 = time.time()
    O.ark_msm(P, S, threads=threads)
    dt = time.time() - t0
    return n / dt, dt


def sweep_row(M, O, KA, torch, curve, lg, peak_gbs, warmup=2, steps=3):
    """One size of the sweep: fresh engine, generated bases (2^15 distinct, replicated), closed-form check, CUDA-event timing."""
    import numpy as np
    n = 1 << lg
    e = M.Engine(curve)
    try:
        e.set_bases_generated(n, seed=SEED_BASES, mode=M.BASES_HARNESS, distinct=min(DISTINCT, n))
        st = torch.cuda.Stream()
        e.set_stream(st.cuda_stream)
        sc = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
        out = torch.zeros(144, dtype=torch.uint8, device="cuda")
        e.generate_scalars(sc.data_ptr(), n, 5)
        e.msm_device(sc.data_ptr(), n, out.data_ptr())
        e.sync()
        r_mod, t, rows = O.constants(curve)["r"], 0, 1 << 24
        for a0 in range(0, n, rows):          # the scalars come back in pieces: 2^28 of them are 8.6 GB
            k = min(rows, n - a0)
            i = np.arange(a0, a0 + k, dtype=np.uint64)
            logs = KA.splitmix64(np.uint64(SEED_BASES) + (i % np.uint64(min(DISTINCT, n)))) | np.uint64(1)
            t = (t + KA.dot_mod_r(sc[a0 * 32:(a0 + k) * 32].cpu().numpy().reshape(k, 32), logs, r_mod)) % r_mod
        ok = bool(np.array_equal(O.jac_to_affine(out.cpu().numpy(), curve)[0], KA.expected_affine(O, curve, t)))
        for _ in range(warmup):
            e.msm_device(sc.data_ptr(), n, out.data_ptr())
        e.sync()
        e.reset_stage_stats()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(steps):
            e.msm_device(sc.data_ptr(), n, out.data_ptr())
        e1.record(st)
        e.sync()
        ms = e0.elapsed_time(e1) / steps
        stg, k = e.stage_ms_sum()
        acc = stg["accumulate"] / max(k, 1)
        info = e.info()
        return {"log_n": lg, "ms": round(ms, 3), "points_per_s": round(n / ms * 1e3), "bit_exact": ok, "window_bits": info["window_bits"], "tables": info["tables"],
                "folded_windows": info["folded_windows"], "buckets": info["buckets"], "accumulate_ms": round(acc, 3),
                "roofline_frac": n * 128 / (acc / 1e3) / 1e9 / peak_gbs}
    finally:
        e.close()
        torch.cuda.empty_cache()


def fastest_reference_entry(ms_ours_total):
    """The reference's fastest GPU entry re-timed on this pool's B200 (committed measurement, profiles/r02_reference_entry.json:
    scripts/retime_reference_entry.py, the entry built by oracle/build_ref_entry.py) beside this run's 4-batch contest call."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "r02_reference_entry.json")))
        out = dict(d)
        if d.get("ms_total") and ms_ours_total:
            out["ratio"] = d["ms_total"] / ms_ours_total
            out["ratio_meaning"] = "reference entry's 4 x 2^26 MSMRun time / this run's contest_call.pageable_ms_total (same call, same ABI)"
        return out
    except Exception:
        return None


def ark_threads(O, n):
    c = O.window_bits(n)
    windows = (253 + c - 1) // c
    return min(os.cpu_count() or 1, windows)


def run_reference(args):
    """The reference's CPU path (arkworks VariableBaseMSM) -- here its C port, oracle/ark_msm_oracle.c -- on host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    O.build()
    lg = args.log_n
    # ONE sample size whatever the host: 2^22 points per step (~4 s on 16 cores: 25 steps in 100 s; arkworks' window rule gives
    # c = 17, 15 windows = 15 threads there, against c = 19, 14 windows at 2^26).  Only a host so slow that K + W such steps would
    # take more than 7 minutes gets the smaller fixed sample 2^20, and the line says so.
    ls = min(REF_SAMPLE_LOG_N, lg)
    cpu_oracle_rate(O, 16)
    note = ""
    _, t_first = cpu_oracle_rate(O, ls)              # first warm-up step, also the guard: K + W steps must end within minutes
    if t_first * (args.steps + args.warmup) > 420.0 and ls > 20:
        note = f" (a 2^{ls} step takes {t_first:.1f} s on this host: sample reduced to 2^20 so that {args.steps} + {args.warmup} steps end within minutes)"
        ls = 20
    for _ in range(max(args.warmup - 1, 0) if not note else args.warmup):
        cpu_oracle_rate(O, ls)
    pure = 0.0                      # the MSM call alone is timed; input generation is outside
    for _ in range(args.steps):
        _, dt = cpu_oracle_rate(O, ls)
        pure += dt
    ms = pure / args.steps * 1e3
    value = (1 << ls) / (ms / 1e3)
    cores = ark_threads(O, 1 << ls)
    c_s, c_f = O.window_bits(1 << ls), O.window_bits(1 << lg)
    sample = (f"2^{ls} points per step, fixed (harness-shaped: 2^10 distinct bases replicated, uniform scalars), arkworks window rule c={c_s} "
              f"({-(-253 // c_s)} windows = threads); at 2^{lg} the rule gives c={c_f} ({-(-253 // c_f)} windows)" + note)
    # one 2^lg MSM at this rate, corrected for the window count: per point the work is one mixed add per window
    extrap_ms = (1 << lg) / value * 1e3 * (-(-253 // c_f)) / (-(-253 // c_s))
    line = {"impl": "reference", "metric": METRIC.format(lg=lg), "value": value, "unit": "points/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64 (377-bit modular integers as 6 x u64 Montgomery limbs, unsigned __int128 products)",
            "data": "synthetic", "config": {"workload": WORKLOAD.format(lg=lg), "log_n": lg, "cpu_sample_log_n": ls},
            f"cpu_2pow{lg}_extrapolated_ms": extrap_ms,
            "cpu_baseline": {"value": value, "unit": "points/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--log-n", type=int, default=26)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--window-bits", type=int, default=0)
    ap.add_argument("--tables", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-pipeline", action="store_true", help="device-resident steps do not overlap (each MSM's tail ends before the next MSM starts)")
    ap.add_argument("--no-peer-reduce", action="store_true", help="N > 1: every rank reduces all of its own buckets (round-1 path)")
    ap.add_argument("--no-sweep", action="store_true", help="skip the size sweep and the BLS12-381 block of the N = 1 line")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    M = g.load_package()
    import importlib
    SH = importlib.import_module(g.PKG_NAME + ".sharded")
    from oracle import oracle as O           # checker + cpu_baseline leg only
    import known_answer as KA

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this engine has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL's log goes to stdout by default and this program prints exactly one JSON line there: route the log to stderr
        # instead of silencing it (the communicator's rank count is read from its INIT lines)
        # (at level VERSION / WARN, which this image presets, NCCL printf()s a bare version banner to stdout; at INFO the banner
        # is a log line like the rest.)  An environment that already chose a log file is left alone.
        if "NCCL_DEBUG_FILE" not in os.environ:
            os.environ["NCCL_DEBUG"] = "INFO"
            os.environ["NCCL_DEBUG_SUBSYS"] = "INIT"
            os.environ["NCCL_DEBUG_FILE"] = "/dev/stderr"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lg = args.log_n
    n = 1 << lg
    lo, hi = SH.shard_bounds(n, world, rank)
    n_local = hi - lo
    if n_local % DISTINCT and n_local > DISTINCT:
        raise SystemExit("shard size must be a multiple of the replication period")

    # ---- init (untimed, like multi_scalar_mult_init): bases = 2^15 random multiples of G replicated to n, sharded
    t_init = time.time()
    eng = M.Engine(M.BLS12_377, device=local, window_bits=args.window_bits, tables=args.tables)
    distinct = min(DISTINCT, n_local)
    eng.set_bases_generated(n_local, seed=SEED_BASES, mode=M.BASES_HARNESS, distinct=distinct)
    backend = SH.CudaShardBackend(eng)
    sh = SH.ShardedMsm(backend)
    stream = backend.stream
    scal = torch.empty(n_local * 32, dtype=torch.uint8, device="cuda")
    eng.generate_scalars(scal.data_ptr(), n_local, SEED_SCALARS + rank)
    eng.sync()
    t_init = time.time() - t_init

    # ---- correctness gate (untimed): closed form vs the CPU oracle
    out = sh.msm_device(scal)
    eng.sync()
    host_scal = torch.empty(n_local * 32, dtype=torch.uint8, pin_memory=True)
    host_scal.copy_(scal)
    torch.cuda.synchronize()
    r_mod = O.constants(0)["r"]
    logs = KA.base_logs(n_local, SEED_BASES, 0, distinct)
    t_local = KA.dot_mod_r(host_scal.numpy().reshape(n_local, 32), logs, r_mod)
    if world > 1:
        parts = [None] * world
        dist.all_gather_object(parts, t_local)
        t_all = sum(parts) % r_mod
    else:
        t_all = t_local
    exp = KA.expected_affine(O, 0, t_all)
    got = O.jac_to_affine(out.cpu().numpy())[0]
    if not np.array_equal(got, exp):
        raise SystemExit(f"rank {rank}: MSM result differs from the oracle's closed form -- refusing to time a wrong kernel")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        eng.reset_stage_stats()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler = ClockSampler(local)
        sampler.start()
        with torch.cuda.stream(stream):
            e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(backend.result_stream)     # the last MSM's result is complete on the stream that carries it
        barrier()
        clocks = sampler.result()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) / steps, clocks

    # ---- device-resident timing.  N > 1: the bucket reduce-scatter over NVLink (rank g reduces slice g of every rank's buckets;
    # sharded.py attach_peers) unless --no-peer-reduce; its result must equal the plain path's before it is timed
    peer = None
    if world > 1 and not args.no_peer_reduce:
        try:
            if sh.attach_peers(chunks=1):
                out_p = sh.msm_device(scal)
                eng.sync()
                if not np.array_equal(O.jac_to_affine(out_p.cpu().numpy())[0], exp):
                    raise SystemExit(f"rank {rank}: bucket reduce-scatter result differs from the oracle")
                peer = eng.peer_info()
        except M.MsmError as ex:
            peer = {"skipped": str(ex)[:160]}
    # first strictly one MSM after another (the stage times and `isolated_ms_per_msm` come from this pass), then -- the headline --
    # pipelined: the latency-bound tail of one MSM under the planning kernels of the next
    ms_iso, clocks = timed(lambda: sh.msm_device(scal), args.steps, args.warmup)
    stage_sum, covered = eng.stage_ms_sum()
    stage_avg = {k: v / max(covered, 1) for k, v in stage_sum.items()}
    ms_dev = ms_iso
    if not args.no_pipeline:
        backend.set_pipelined(True)
        ms_dev, clocks = timed(lambda: sh.msm_device(scal), args.steps, args.warmup)
        eng.sync()
        if not np.array_equal(O.jac_to_affine(sh.out.cpu().numpy())[0], exp):
            raise SystemExit(f"rank {rank}: pipelined MSM result differs from the oracle")
    if peer and "world" in peer:
        peer = eng.peer_info()              # with the last timed MSM's wait / exchange times
    info = eng.info()
    launches_per_msm = info["launches"] + (1 if world > 1 else 0)      # + the combine kernel after the all-gather

    # ---- end to end: host scalars through the harness call.  Headline = PAGEABLE host memory, which is what the harness hands
    # over (a Rust Vec); the library stages it through pinned slots itself (csrc/msm.cu h2d).  The same from pinned memory is
    # reported beside it.
    e2e = None
    if not args.no_e2e:
        if world > 1 and sh.peers:
            sh.attach_peers(chunks=3 if n_local >= (1 << 24) else 2 if n_local >= (1 << 22) else 1)      # = the library's default split of a host-memory call      # the chunked upload keeps one bucket array per chunk (msm.cu upload_and_run)
        pageable_scal = torch.empty(n_local * 32, dtype=torch.uint8)
        pageable_scal.copy_(host_scal)
        res_holder = {}
        if world == 1:
            hctx = M.MultiScalarMultContext.from_engine(eng)

            def make_step(buf):
                arr = buf.numpy()
                return lambda: res_holder.__setitem__("r", M.multi_scalar_mult(hctx, None, arr))      # mult_pippenger_inf: H2D + MSM + D2H, synchronous
        else:
            def make_step(buf):
                return lambda: res_holder.__setitem__("r", sh.msm(buf))

        def time_e2e(step):
            for _ in range(2):
                step()
            barrier()
            t0 = time.perf_counter()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(args.steps):
                step()
            e1.record(stream)
            barrier()
            wall_ms = (time.perf_counter() - t0) * 1e3 / args.steps
            t = torch.tensor([max(e0.elapsed_time(e1) / args.steps, wall_ms if world == 1 else 0.0)], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            r = res_holder["r"]
            r_np = r.numpy() if hasattr(r, "numpy") else np.asarray(r)
            if not np.array_equal(O.jac_to_affine(r_np.reshape(-1, 144)[0])[0], exp):
                raise SystemExit("e2e result differs from the oracle")
            return float(t.item())

        ms_pageable = time_e2e(make_step(pageable_scal))
        ms_pinned = time_e2e(make_step(host_scal))
        # N = 1 is the harness call, whose scalars are a pageable Vec: that is the headline.  N > 1 has no harness counterpart; the
        # sharded API documents pinned per-rank buffers (N processes staging pageable memory through N copy pools share one
        # host's cores), so its headline is the pinned number and the pageable one is reported beside it.
        ms_e2e, other, other_ms = (ms_pageable, "pinned", ms_pinned) if world == 1 else (ms_pinned, "pageable", ms_pageable)
        e2e = {"value": n / (ms_e2e / 1e3), "unit": "points/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": n * 32, "d2h_bytes_per_step": 144 * world,
               "host_memory": "pageable (staged through pinned slots by the library)" if world == 1 else "pinned, one buffer per rank",
               other: {"value": n / (other_ms / 1e3), "ms_per_step": other_ms},
               "api": "multi_scalar_mult -> mult_pippenger_inf (C ABI)" if world == 1 else "ShardedMsm.msm (host scalars per rank, bucket reduce-scatter over NVLink + NCCL all-gather of partials)" if sh.peers else "ShardedMsm.msm (host scalars per rank, NCCL all-gather of partials)"}
        del pageable_scal

    # ---- the contest's timed unit (P1A/*/benches/msm.rs: one multi_scalar_mult call with 4 batches of scalars in host
    # memory), reported beside the headline for comparison with the published A40 numbers; N = 1 only
    contest = None
    if world == 1 and not args.no_e2e and lg <= 26:
        try:
            host4 = torch.empty(4 * n_local * 32, dtype=torch.uint8, pin_memory=True)
            for b in range(4):
                eng.generate_scalars(scal.data_ptr(), n_local, SEED_SCALARS + 100 + b)
                eng.sync()
                host4[b * n_local * 32:(b + 1) * n_local * 32].copy_(scal)
            torch.cuda.synchronize()
            sc4 = host4.numpy()
            M.multi_scalar_mult(hctx, None, sc4)
            t0 = time.perf_counter()
            r4 = M.multi_scalar_mult(hctx, None, sc4)
            dt = time.perf_counter() - t0
            contest = {"batches": 4, "ms_total": dt * 1e3, "ms_per_msm": dt * 1e3 / 4, "results": int(r4.shape[0]), "scalars": "pinned host memory",
                       "published_A40_ms_total": "2200-2300 (combined-top-solutions README; other hardware)"}
            try:                         # the same call with the scalars in pageable memory, which is what a Rust Vec is
                pageable = np.empty_like(sc4)
                pageable[:] = sc4
                M.multi_scalar_mult(hctx, None, pageable)
                t0 = time.perf_counter()
                r4p = M.multi_scalar_mult(hctx, None, pageable)
                contest["pageable_ms_total"] = (time.perf_counter() - t0) * 1e3
                contest["pageable_same_result"] = bool(np.array_equal(r4p, r4))
                del pageable
            except Exception as ex:
                contest["pageable_skipped"] = str(ex)[:120]
            del host4
        except Exception as ex:          # pinned 8 GiB may not be available everywhere; the headline does not depend on it
            contest = {"skipped": str(ex)[:120]}

    # ---- arkworks' own call shape (VariableBaseMSM::multi_scalar_mul: bases AND scalars from host memory, no context, no
    # table precompute), reported beside the headline at the headline size; N = 1 only
    one_call = None
    if world == 1 and not args.no_e2e:
        try:
            m = n_local
            hb = eng.get_bases(0, m)                                   # uint8[m, 96]: x, y in arkworks' Montgomery form
            hs = host_scal.numpy()[:m * 32]
            V = M.VariableBaseMSM
            r1 = V.multi_scalar_mul(hb, hs, stride=96)                 # first call: allocates the one-table plan
            t0 = time.perf_counter()
            for _ in range(2):
                r1 = V.multi_scalar_mul(hb, hs, stride=96)
            dt = (time.perf_counter() - t0) / 2
            exp1 = KA.expected_affine(O, 0, KA.dot_mod_r(hs.reshape(m, 32), logs[:m], r_mod))
            ok = bool(np.array_equal(O.jac_to_affine(r1)[0], exp1))
            V.release()
            one_call = {"log_n": m.bit_length() - 1, "ms": dt * 1e3, "points_per_s": m / dt, "bit_exact": ok,
                        "h2d_bytes": m * (96 + 32), "api": "VariableBaseMSM.multi_scalar_mul -> b200msm_variable_base_msm (pageable host bases + pinned scalars)"}
            del hb
        except Exception as ex:
            one_call = {"skipped": str(ex)[:120]}

    # ---- BASELINE configs 3 and 5 on the same box (N = 1): size sweep 2^20 .. 2^28 with the auto-tuned window plan, and BLS12-381
    # at 2^24, every row checked against the closed form before it is timed.  The headline engine is released first: 2^28
    # needs the whole HBM (3 tables of 38.7 GB).
    sweep, second_curve = None, None
    if world == 1 and not args.no_sweep and lg == 26:
        stage_keep = dict(stage_avg)
        eng.close()
        del scal
        torch.cuda.empty_cache()
        peak_gbs, _ = measured_peak_gbs()
        sweep = []
        for curve, slg in ((0, 20), (0, 22), (0, 24), (0, 26), (0, 28), (1, 24)):
            if curve == 0 and slg == lg:
                row = {"log_n": lg, "ms": round(ms_dev, 3), "points_per_s": round(n / ms_dev * 1e3), "bit_exact": True, "window_bits": info["window_bits"], "tables": info["tables"],
                       "folded_windows": info["folded_windows"], "buckets": info["buckets"], "accumulate_ms": round(stage_keep["accumulate"], 3),
                       "roofline_frac": n * 128 / (stage_keep["accumulate"] / 1e3) / 1e9 / peak_gbs}
            else:
                try:
                    row = sweep_row(M, O, KA, torch, curve, slg, peak_gbs)
                except Exception as ex:
                    row = {"log_n": slg, "error": str(ex)[:160]}
            if curve == 1:
                second_curve = dict(row, curve="BLS12-381", bucket_arithmetic="XYZZ (8 M + 2 S mixed add; no twisted-Edwards model)")
            else:
                sweep.append(row)

    # ---- CPU baseline beside it (rank 0, N = 1): bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        O.build()
        ls = min(REF_SAMPLE_LOG_N, lg)
        cpu_oracle_rate(O, 16)
        rate, dt = cpu_oracle_rate(O, ls)
        c_s, c_f = O.window_bits(1 << ls), O.window_bits(1 << lg)
        cpu = {"value": rate, "unit": "points/s", "cores": ark_threads(O, 1 << ls), "kind": "port",
               "sample": f"one 2^{ls}-point MSM ({dt:.1f} s), the sample size of --impl reference; harness-shaped inputs, arkworks window rule c={c_s}; host has {os.cpu_count()} cores",
               f"cpu_2pow{lg}_extrapolated_ms": (1 << lg) / rate * 1e3 * (-(-253 // c_f)) / (-(-253 // c_s))}

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        acc_ms = stage_avg["accumulate"]
        alg_bytes = n_local * 128
        achieved = alg_bytes / (acc_ms / 1e3) / 1e9
        # integer-pipe view: mixed adds per launch x IMAD.WIDE per mixed add (twisted-Edwards unified add: 7 multiplies x 302) vs the
        # measured IMAD.WIDE.U32 rate: 32 lanes / clk / SM -- the fmaheavy pipe is 89 % busy when the multiplier
        # microbenchmark sustains 378 IMAD.WIDE per 13.2 SM-clocks (profiles/r01_mul_bench.txt)
        madds = info["windows"] * n_local
        imad_per_madd = 7 * 302                  # 13 x 29-bit limbs, one-level Karatsuba + 14 reduction rounds (fp29.cuh)
        sm_mhz = clocks["sm_mhz"] or 1965
        int_peak = 32.0 * 148 * sm_mhz * 1e6
        int_ach = madds * imad_per_madd / (acc_ms / 1e3)
        traffic = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", TRAFFIC_FILE)))["k_accumulate"]
            if tr["log_n"] == lg and tr["n_gpus"] == world:
                traffic = tr["traffic_bytes"]          # dram read + write of one launch, from the committed ncu --set full capture
        except Exception:
            pass
        line = {
            "metric": METRIC.format(lg=lg), "value": n / (ms_dev / 1e3), "unit": "points/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u32 (377-bit modular integers: 13 x 29-bit limbs in the accumulate kernel, 14 x 28-bit elsewhere; IMAD.WIDE into u64 columns)", "data": "synthetic",
            "config": {"workload": WORKLOAD.format(lg=lg),
                       "log_n": lg, "points_per_gpu": n_local, "parallelism": f"point-shard x{world}" if world > 1 else "single GPU",
                       "window_bits": info["window_bits"], "windows": info["windows"], "tables": info["tables"], "folded_windows": info["folded_windows"],
                       "buckets": info["buckets"], "l2": "inputs larger than L2 (scalars %.1f GB, base tables %.1f GB per GPU)" % (n_local * 32 / 1e9, info["tables"] * n_local * M.ENTRY_BYTES[M.BLS12_377] / 1e9),
                       "isolated_ms_per_msm": ms_iso,
                       "steps_overlap": ("consecutive MSMs are pipelined on one context: the ~3 ms latency-bound tail of one (upper reduction levels, finalize"
                                         + (", all-gather, combine" if world > 1 else "") + ") runs beside the digit / scatter kernels of the next; --no-pipeline times them strictly one after another")
                                        if not args.no_pipeline else "none (--no-pipeline)",
                       "init_s": round(t_init, 2), "checked": "bit-exact vs CPU oracle (closed form over known discrete logs) before timing"},
            "stages_ms": {k: round(v, 3) for k, v in stage_avg.items()},
            "roofline": {"bound": "hbm", "kernel": "k_accumulate", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": peak_src, "ms_per_launch": acc_ms, "algorithmic_bytes": alg_bytes,
                         "note": "integer-pipe bound by construction; see int_pipe"},
            "int_pipe": {"imad_wide_per_launch": madds * imad_per_madd, "achieved_per_s": int_ach, "peak_per_s": int_peak, "frac": int_ach / int_peak,
                         "peak_source": "32 IMAD.WIDE lanes/clk/SM measured (profiles/r01_mul_bench.txt) x 148 SMs x sampled SM clock"},
            "cpu_baseline": cpu, "e2e": e2e, "contest_call": contest, "one_call": one_call, "gpu_launches": launches_per_msm * args.steps, "clocks": clocks,
        }
        if sweep is not None:
            line["sweep"] = sweep
            line["bls12_381_2pow24"] = second_curve
        if world == 1:
            line["fastest_entry"] = fastest_reference_entry((contest or {}).get("pageable_ms_total") or (contest or {}).get("ms_total"))
        if peer:
            line["bucket_reduce_scatter"] = dict(peer, note="rank g reduces slice g of every rank's buckets, read over NVLink inside k_reduce_slices; bytes = one rank, one MSM",
                                                 nvlink_GBps_in_exchange_kernel=(peer.get("nvlink_bytes", 0) / 1e9) / (peer["last_exchange_us"] / 1e6) if peer.get("last_exchange_us") else None,
                                                 nvlink_peer_copy_peak_GBps=770.0)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- AdaPGM iterations/s on the dense fp64 lasso (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host cores

A step is ONE AdaPGM iteration (src/AdaProx.jl:334-362: value and gradient of
the least-squares term, the stepsize rule, the prox step) on the planted lasso
of lasso/runme.jl:40-77 generated on the device.  N = 1 runs configs[3] at its
full size, 65536 x 131072 fp64 (68.7 GB, far larger than the 126 MB L2, so no
L2 flush is needed between iterations).  N > 1 row-shards the same instance
(strong scaling): per iteration each rank sweeps its shard and the A'r partials
(n + 2 doubles) are all-reduced -- inside the sweep kernel over NVLink peer
memory by default, with ncclAllReduce under --nccl.

The default kernel is the single-sweep fused kernel (A is read ONCE per
iteration: g = A'(Ax - b) per row block while the rows are in shared memory);
--two-pass times the two-sweep persistent kernel (A*x, then A'r) for the A/B.
`roofline.achieved` counts the bytes the kernel that ran has to move.

`value` comes from CUDA events recorded by the library on the stream its kernels
run on, around a solve of exactly K iterations with everything resident in HBM
(the solve's prologue -- one more gradient evaluation -- is inside the timed
region, so the figure is slightly pessimistic).  `e2e` is the same solve timed
from the caller's side of the C ABI with host buffers: x0 copied in from pinned
host memory, x and the per-iteration records copied back.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

M_FULL, N_FULL = 65536, 131072
METRIC = "AdaPGM iters/sec on 65536x131072 fp64 lasso"


def b_iter_bytes(m_loc, n, passes=2):
    """Algorithmic bytes of one AdaPGM lasso iteration.  passes = 2 is SURVEY.md section 8d's figure (A read once
    for A*x and once for A'*r, the reference's two oracle calls) plus the vector traffic; passes = 1 is what the
    single-pass fused kernel (solver_fused.cuh) has to move: A once."""
    return passes * 8 * m_loc * n + 8 * (3 * m_loc + 8 * n)


NCU_SUMMARY = {1: "r01_ncu_k_adapgm_fused.json", 2: "r01_ncu_k_primal_dual_ring.json"}


def ncu_traffic_per_eval(m, n, passes):
    """DRAM bytes (read + write) per gradient evaluation from the committed `ncu --set full` capture of the
    persistent kernel that ran, at the full size; None for other sizes or when no capture is committed."""
    p = os.path.join(ROOT, "profiles", NCU_SUMMARY.get(passes, ""))
    if (m, n) != (M_FULL, N_FULL) or not os.path.isfile(p):
        return None, None
    with open(p) as f:
        return float(json.load(f)["_derived"]["dram_bytes_per_gradient_eval"]), os.path.relpath(p, ROOT)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region: NVML polled every 5 ms from a thread (the timed region
    of a multi-GPU run lasts only tens of milliseconds), nvidia-smi at 100 ms as the fallback."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None
        self.nv, self.samples, self.stop_flag, self.t = None, [], threading.Event(), None

    def _nvml_loop(self):
        nv, hd = self.nv, self.handle
        while not self.stop_flag.is_set():
            try:
                self.samples.append((nv.nvmlDeviceGetClockInfo(hd, nv.NVML_CLOCK_SM), nv.nvmlDeviceGetCurrentClocksEventReasons(hd)))
            except Exception:
                break
            time.sleep(0.005)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.handle = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(self.handle, nv.NVML_CLOCK_SM)
            self.nv = nv
            self.t = threading.Thread(target=self._nvml_loop, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nv = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def stop(self):
        if self.nv is not None:
            self.stop_flag.set()
            self.t.join(timeout=2)
            nv = self.nv
            names = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                     "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
            sm = [s[0] for s in self.samples]
            reasons = sorted(k for k, bit in names.items() if any(s[1] & bit for s in self.samples))
            return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.max_mhz, "samples": len(sm),
                    "reasons": reasons, "source": "NVML, 5 ms period"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            parts = [s.strip() for s in l.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi, 100 ms period"}


# ----------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm (oracle port; Julia is not installed) on the host cores
# ----------------------------------------------------------------------------------------------
def cpu_reference(m, n, steps, warmup, rows_sample=None, seed=0):
    """AdaPGM (OurRule) on a row slab of the same n, Fortran order like Julia's
    Matrix, numpy -> OpenBLAS dgemv on all host threads.  Returns (it/s scaled to
    the full m, description)."""
    from oracle import adaprox_oracle as O
    try:
        from threadpoolctl import threadpool_info
        thr = max([d.get("num_threads", 1) for d in threadpool_info()] or [os.cpu_count()])
    except Exception:
        thr = os.cpu_count()
    if rows_sample is None:
        rows_sample = max(64, min(m, int(1.0e9 // (8 * n))))       # ~1 GB slab
    rng = np.random.default_rng(seed)
    A = np.asfortranarray(rng.random((rows_sample, n)) * 2.0 - 1.0)
    A *= 1.0 / np.sqrt(n)
    b = rng.random(rows_sample)
    f, g = O.LinearLeastSquares(A, b), O.NormL1(1.0)
    v = np.ones(n) / np.sqrt(n)
    for _ in range(3):
        w = A.T @ (A @ v); lf = float(np.linalg.norm(w)); v = w / lf
    rule = O.OurRule(gamma=1.0 / lf)
    if warmup > 0:
        O.adaptive_proxgrad(np.zeros(n), f=f, g=g, rule=rule, tol=0.0, maxit=warmup)
    t0 = time.perf_counter()
    O.adaptive_proxgrad(np.zeros(n), f=f, g=g, rule=rule, tol=0.0, maxit=steps)
    dt = time.perf_counter() - t0
    per_iter_full = dt / steps * (m / rows_sample)
    desc = (f"oracle port of src/AdaProx.jl:312-364 + lasso/runme.jl:16-27 (numpy/OpenBLAS dgemv, {thr} threads), "
            f"{steps} iterations on a {rows_sample}x{n} row slab ({rows_sample * n * 8 / 1e9:.2f} GB, Fortran order), "
            f"time scaled by m/rows = {m / rows_sample:.1f} to the full {m}x{n}")
    return 1.0 / per_iter_full, thr, desc, dt / steps * 1e3


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    m, n = args.m, args.n
    val, thr, desc, ms_sample = cpu_reference(m, n, args.steps, args.warmup, args.cpu_rows)
    out = {
        "impl": "reference", "metric": METRIC if (m, n) == (M_FULL, N_FULL) else f"AdaPGM iters/sec on {m}x{n} fp64 lasso",
        "value": val, "unit": "it/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 / val, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"configs[3]: dense lasso AdaPGM {m}x{n} fp64, OurRule, CPU", "m": m, "n": n},
        "cpu_baseline": {"value": val, "unit": "it/s", "cores": thr, "kind": "port", "sample": desc},
        "e2e": {"value": val, "unit": "it/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import adaprox_b200 as AdaProx
    import ctypes as C
    from adaprox_b200 import _lib as L

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch N > 1 with torch.distributed.run (one process per GPU)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    dev = AdaProx.Device(local)
    AdaProx.set_default_device(dev)
    if world > 1:
        AdaProx.sharding.attach_communicator(dev, dist)
        if not args.nccl:
            AdaProx.sharding.attach_p2p(dev, args.n, dist)       # all-reduce inside the sweep kernel over NVLink peer memory

    m, n = args.m, args.n
    row0, rows = AdaProx.sharding.shard_rows(m, world, rank)
    t_gen = time.perf_counter()
    P = AdaProx.generate_planted_lasso(m, n, pfactor=5, seed=0, lam=1.0, power_iters=args.power_iters, row0=row0, rows=rows, dev=dev)
    t_gen = time.perf_counter() - t_gen
    f = AdaProx.LinearLeastSquares(P["A"], P["b"])
    g = AdaProx.NormL1(1.0)
    gamma0 = 1.0 / P["Lf"]

    # pinned host buffers for the C-ABI call
    x0_t = torch.zeros(n, dtype=torch.float64).pin_memory()
    xo_t = torch.zeros(n, dtype=torch.float64).pin_memory()
    x0 = x0_t.numpy(); xo = xo_t.numpy()

    def solve(maxit, tol, records):
        p = f._problem(n)
        p.g = g._desc()
        o = AdaProx.core._opts(tol, maxit)
        AdaProx.OurRule(gamma=gamma0)._fill(o)
        o.solver = L.S_ADAPTIVE_PROXGRAD
        o.want_objective = 1 if records else 0
        o.max_records = maxit if records else 0
        recs = (L.Record * max(maxit, 1))() if records else None
        res = L.Result()
        dev.check(dev.lib.adaprox_solve(dev.h, C.byref(p), C.byref(o), x0.ctypes.data_as(L.c_dp), None,
                                        xo.ctypes.data_as(L.c_dp), None, recs, C.byref(res)))
        return res, recs

    def sync_all():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(v):
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    W, K = max(args.warmup, 3), args.steps
    solve(W, 0.0, False)                                       # warm-up iterations (untimed)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    sync_all()
    t0 = time.perf_counter()
    res, recs = solve(K, 0.0, True)                            # exactly K iterations, blocking C-ABI call
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    sync_all()
    clocks = sampler.stop() if rank == 0 else None
    dev_ms = max_over_ranks(res.solve_ms)
    e2e_s = max_over_ranks(t1 - t0)
    launches = int(res.kernel_launches)
    assert res.iters == K, (res.iters, K)
    last = recs[K - 1]

    # per-kernel-family timing (profile breakdown; the shares the ncu launch list must agree with)
    ms_n = max_over_ranks(P["A"].time_kernel(0, reps=3))
    ms_t = max_over_ranks(P["A"].time_kernel(1, reps=3))

    extra = {}
    if args.to_tol > 0:
        sync_all()
        t0 = time.perf_counter()
        r2, _ = solve(args.tol_maxit, args.to_tol, False)
        torch.cuda.synchronize()
        tt = max_over_ranks(time.perf_counter() - t0)
        extra = {"time_to_tol": {"tol": args.to_tol, "seconds": tt, "device_seconds": max_over_ranks(r2.solve_ms) / 1e3,
                                 "iterations": int(r2.iters), "converged": bool(r2.flags & 1),
                                 "final_norm_res": r2.final_norm_res}}

    if rank == 0:
        peak, peak_src = peaks()
        passes = int(res.matrix_passes)                            # 1: single-pass fused kernel, 2: A*x then A'r
        bytes_iter_rank = b_iter_bytes(rows, n, passes)
        # the timed launch(es) perform K iterations plus the solver's prologue (one more gradient evaluation, :327-332)
        bytes_launch = bytes_iter_rank * K + passes * 8 * rows * n
        achieved = bytes_launch / (dev_ms * 1e-3) / 1e9            # per-GPU GB/s actually required of HBM
        per_eval, per_eval_src = ncu_traffic_per_eval(m, n, passes) if world == 1 else (None, None)
        if world > 1 and passes == 1:
            kernel = "k_adapgm_fused in sweep-only mode (one sweep of the row shard per iteration)"
        elif world > 1:
            kernel = "k_sh_A + k_sh_C (split-phase GEMV kernels)"
        elif passes == 1:
            kernel = ("k_adapgm_fused (persistent cooperative cluster kernel: the whole solve is one launch; "
                      "A'(Ax-b) in ONE sweep over A per iteration)")
        else:
            kernel = "k_primal_dual<false> (persistent cooperative kernel: the whole solve is one launch)"
        out = {
            "metric": METRIC if (m, n) == (M_FULL, N_FULL) else f"AdaPGM iters/sec on {m}x{n} fp64 lasso",
            "value": K / (dev_ms * 1e-3), "unit": "it/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"configs[3]: dense lasso AdaPGM {m}x{n} fp64 (planted instance of lasso/runme.jl:40-77), "
                                   f"OurRule gamma0=1/Lf, row-sharded over {world} GPU(s)",
                       "m": m, "n": n, "rows_per_gpu": rows, "rule": "OurRule", "lambda": 1.0,
                       "l2": "inputs larger than L2 (matrix shard %.1f GB vs 126 MB), no flush needed" % (rows * n * 8 / 1e9),
                       "timing": "library CUDA events on the kernels' stream around one solve of K iterations; max over ranks",
                       "generation_s": round(t_gen, 2), "gamma0": gamma0},
            "achieved_hbm_gbs_per_gpu": bytes_iter_rank * K / (dev_ms * 1e-3) / 1e9,
            "matrix_passes_per_iteration": passes,
            "two_pass_equivalent_gbs_per_gpu": b_iter_bytes(rows, n, 2) * K / (dev_ms * 1e-3) / 1e9,
            "two_pass_equivalent_note": "SURVEY 8d's B_iter (A counted twice, as the reference's two oracle calls read it) x it/s; "
                                        "equals achieved_hbm_gbs_per_gpu when matrix_passes_per_iteration = 2; with the single-pass "
                                        "kernel it is a throughput equivalent, NOT a bandwidth (roofline.achieved counts A once)",
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": (per_eval * (K + 1)) if per_eval else None, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": bytes_launch,
                         "launch": "one persistent launch = K iterations + prologue = K+1 gradient evaluations" if world == 1
                                   else ("K+1 gradient evaluations as 3 launches each (fused sweep + all-reduce, 2 small kernels)" if passes == 1
                                         else "K+1 gradient evaluations as 6 launches + 1 all-reduce each"),
                         "traffic_source": ("ncu dram__bytes_read.sum + dram__bytes_write.sum per gradient evaluation x (K+1), "
                                            + per_eval_src) if per_eval else None,
                         "kernel": kernel, "matrix_passes": passes,
                         "algorithmic_bytes_per_iteration_per_gpu": bytes_iter_rank,
                         "gemv_n_ms": ms_n, "gemv_t_ms": ms_t,
                         "gemv_n_gbs": rows * n * 8 / (ms_n * 1e-3) / 1e9, "gemv_t_gbs": rows * n * 8 / (ms_t * 1e-3) / 1e9},
            "e2e": {"value": K / e2e_s, "unit": "it/s", "h2d_bytes_per_step": n * 8 / K,
                    "d2h_bytes_per_step": (n * 8 + K * C.sizeof(L.Record) + C.sizeof(L.Result)) / K,
                    "call": "one blocking adaprox_solve (C ABI) of K iterations: x0 from pinned host memory, x and K records copied back"},
            "gpu_launches": launches,
            "collective": {0: "none", 1: "ncclAllReduce of n+2 doubles per iteration",
                           2: "all-reduce inside the sweep kernel over NVLink peer memory (no NCCL on the data path)"}[int(res.collective)],
            "clocks": clocks,
            "final_record": {"it": int(last.it), "gamma": last.gamma, "norm_res": last.norm_res,
                             "objective": last.f_x + last.g_x, "optimum": P["optimum"]},
        }
        out.update(extra)
        if world == 1 and not args.no_cpu:
            val, thr, desc, _ = cpu_reference(m, n, args.cpu_steps, 1, args.cpu_rows)
            out["cpu_baseline"] = {"value": val, "unit": "it/s", "cores": thr, "kind": "port", "sample": desc}
        print(json.dumps(out), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--m", type=int, default=M_FULL)
    ap.add_argument("--n", type=int, default=N_FULL)
    ap.add_argument("--power-iters", type=int, default=30)
    ap.add_argument("--to-tol", type=float, default=0.0, help="also report the time to reach norm_res <= tol")
    ap.add_argument("--tol-maxit", type=int, default=20000)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--two-pass", action="store_true", help="A/B: force the two-pass kernel (ADAPROX_FUSED=0)")
    ap.add_argument("--nccl", action="store_true", help="A/B (N > 1): ncclAllReduce per iteration instead of the in-kernel all-reduce")
    ap.add_argument("--cpu-steps", type=int, default=20)
    ap.add_argument("--cpu-rows", type=int, default=None)
    args = ap.parse_args()
    if args.two_pass:
        os.environ["ADAPROX_FUSED"] = "0"
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()

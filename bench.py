#!/usr/bin/env python
"""bench.py — the two headline metrics of BASELINE.json on the B200 engine.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun ... bench.py --gpus N ...                  (one rank per GPU, NCCL)

Metric 1 (the JSON line's `value`): Gibbs sweeps/s summed over all chains on configs[2] ("C3": synthetic 20-variable
VHAR(5,22) with stochastic volatility, Dirichlet-Laplace prior, 1,024 chains per GPU; k=20, d=61, n=1000).  One *step*
= one RECORDED Gibbs sweep of every chain (the nine updaters + updateRecords).  `value` is timed on the device (CUDA
events on the engine's stream, max over ranks) with everything resident in HBM; `e2e` is the same metric through the
public API (`VharBayes.fit()`): host buffers in, posterior draws out, copies inside the timed region.  Weak scaling
(1,024 chains per GPU) is the line's `value`; `strong` holds the same measurement with 1,024 chains IN TOTAL
(north_star: "1,024 chains x 10k sweeps ... scales >= 7x at 8 GPUs").

Metric 2 (`forecast_roll`): rolling-window refits per second on configs[3] ("C4": 50-variable VAR(4), NG prior + SV,
4 chains per window, k=50, d=201, n=400), every window a full refit of `--roll-iters` sweeps plus the one-step density
forecast, through the native out-of-sample runner with host buffers; windows are sharded over the ranks.

The `--impl reference` arm times the CPU restatement of the reference's own algorithm (oracle/, reference-faithful
formulation: Kronecker design + dense n x n SV precision) in the reference's own parallel shape (one OpenMP task per
chain, triangular.h:1222), dense kernels through the OpenBLAS bundled with SciPy.  The reference itself cannot be built
in this image (no R / Rcpp / Eigen / Boost), DESIGN.md section 2.
"""
from __future__ import annotations

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

METRIC = "gibbs_sweeps_per_s_all_chains_vhar_sv_m20"
UNIT = "chain-sweeps/s"
CHAINS_PER_GPU = 1024
STRONG_CHAINS_TOTAL = 1024


def workload_config(chains, prior, world):
    """The `config` object of the JSON line: identical for both arms (it names the workload, not the arm)."""
    return {"workload": f"C3: synthetic 20-variable VHAR(5,22)-SV, {prior} prior, {chains} chains per GPU, k=20 d=61 n=1000",
            "chains_total": world * chains,
            "l2": "per-sweep working set (>1 GB of chain state) exceeds the 126 MB L2",
            "timed_path": "recorded sweeps (9 updaters + updateRecords), CUDA graph replay; roofline from a second pass with per-kernel events"}


def build_model(chains, iters, burn, seed, device, record_mask=None, prior="DL", dim=20, T=1022):
    from bvhar_b200 import _spec as sp
    from bvhar_b200 import datasets
    from bvhar_b200.model import VharBayes
    cfg = {"DL": sp.DlConfig, "Horseshoe": sp.HorseshoeConfig, "NG": sp.NgConfig, "SSVS": sp.SsvsConfig,
           "GDP": sp.GdpConfig, "Minnesota": lambda: sp.MinnesotaConfig(is_long=True)}[prior]()
    y = datasets.simulate_vhar_sv(T=T, dim=dim, seed=20240001)
    return VharBayes(y, week=5, month=22, n_chain=chains, n_iter=iters, n_burn=burn, bayes_config=cfg,
                     cov_config=sp.SvConfig(), seed=seed, device=device, record_mask=record_mask)


def algorithmic_flops(k, d, n):
    """SURVEY.md section 8(d), fp64 flops of one sweep of one chain (SV model)."""
    gram = k * n * d * (d + 1)
    chol = k * d ** 3 / 3
    solve = 3 * k * d * d
    rhs = 4 * k * n * d + 2 * n * k * k
    latent = 2 * n * d * k + 2 * n * k * k
    impact = n * k ** 3 / 3 + sum(j ** 3 / 3 + 3 * j * j for j in range(k))
    state = n * k * (7 * 12 + 30)
    return dict(gram=gram, total=gram + chol + solve + rhs + latent + impact + state)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except Exception:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of the G1 launch from the committed `ncu --set full` capture (same
    workload, 1 024 chains), or None."""
    for name in ("r02_gram_traffic.json", "r01_gram_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                return json.load(f)["traffic_bytes_per_launch"]
        except Exception:
            continue
    return None


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def fp64_peaks(local):
    """fp64 denominators (absent from MEASURED_PEAKS.json): DMMA issue peak, DFMA peak and cuBLAS DGEMM 8192^3, measured here."""
    import ctypes as C

    import torch
    from bvhar_b200 import _abi
    L = _abi.load_library()
    pk = C.c_double()
    L.bvhar_cuda_bench_fp64_peak(local, 1, C.byref(pk))
    dmma = pk.value
    L.bvhar_cuda_bench_fp64_peak(local, 0, C.byref(pk))
    dfma = pk.value
    a = torch.randn(8192, 8192, dtype=torch.float64, device="cuda"); b = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
    torch.matmul(a, b); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for _ in range(3):
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b
    return {"dmma_issue_tflops": dmma, "dfma_tflops": dfma, "cublas_dgemm_tflops": 2 * 8192 ** 3 / (best * 1e-3) / 1e12}


# ------------------------------------------------------------------------------------------------ metric 2: forecast_roll
def roll_problem(device, n_windows, chains, iters, burn, first_window=1):
    """C4: synthetic 50-variable VAR(4), NG prior + SV, rolling windows of 404 rows (n = 400), `chains` chains per window.
    Windows [first_window, first_window + n_windows) of the table (window 0 of forecast_roll reuses the original fit)."""
    from bvhar_b200 import _spec as sp
    from bvhar_b200 import datasets
    from bvhar_b200.model import VarBayes
    y = datasets.simulate_var_sv(404 + first_window + n_windows, 50, 4)
    m = VarBayes(y[:404], 4, chains, iters, burn, 1, sp.NgConfig(), sp.SvConfig(), seed=1, device=device)
    x_tot, y_tot = m._design_of(y)
    n0 = m.response_.shape[0]
    rows = list(range(first_window, first_window + n_windows))
    pk = m._pack(x_tot, y_tot, m._seeds(n_windows, chains), win_row0=rows, win_nrow=[n0] * n_windows, record_mask=0,
                 unit_id_base=first_window * chains)
    return m, pk, y


def roll_metric(device, rank, world, n_windows, chains, iters, burn, barrier, allmax):
    """forecast_roll() windows/s: every (window, chain) refitted with `iters` sweeps and forecast one step ahead through
    bvhar_cuda_outforecast_run (host buffers in, forecast draws out).  Rank r takes windows [1 + r W, 1 + (r + 1) W)."""
    from bvhar_b200._engine import CudaFit, outforecast
    m, pk, y = roll_problem(device, n_windows, chains, iters, burn, first_window=1 + rank * n_windows)
    m_w, pk_w, y_w = roll_problem(device, min(n_windows, 8), chains, 6, 2, first_window=1)
    outforecast(pk_w, y_w, 1, 4, m_w._seeds(chains))  # warm-up: module load, memory pools, function attributes
    barrier()
    t0 = time.perf_counter()
    fc, _ = outforecast(pk, y, 1, 4, m._seeds(chains))
    dt = allmax(time.perf_counter() - t0)
    assert np.all(np.isfinite(fc))
    out = {"metric": "forecast_roll_windows_per_s", "value": world * n_windows / dt, "unit": "windows/s", "seconds": dt,
           "windows_total": world * n_windows, "windows_per_gpu": n_windows, "chains_per_window": chains, "sweeps_per_refit": iters,
           "burn_in": burn, "unit_sweeps_per_s": world * n_windows * chains * iters / dt, "scaling": "weak",
           "config": f"C4: {n_windows} rolling windows per GPU x {chains} chains, k=50 d=201 n=400, NG prior + SV, {iters} sweeps per refit "
                     f"({burn} burn-in), 1-step density forecast",
           "path": "bvhar_cuda_outforecast_run, end to end with host buffers (design in, forecast draws out)"}
    if rank == 0:  # roofline of C4's dominant kernel from a per-kernel event pass over the same batch
        f = CudaFit(pk)
        f.step(2, False)
        f.set_profiling(True); f.step(3, False); prof = f.profile(); f.set_profiling(False)
        f.close()
        k, d, n = 50, 201, 400
        units = n_windows * chains
        g_cnt, g_ms = prof.get("gram_S_dmma", (0, 0.0))
        tot = sum(v[1] for v in prof.values())
        if g_cnt:
            ach = units * k * n * d * (d + 1) / (g_ms / g_cnt * 1e-3) / 1e12
            out["roofline"] = {"kernel": "dgemm_tn_kernel (G1 at C4: K = n = 400)", "bound": "tensor", "achieved": ach, "unit": "TFLOP/s",
                               "avg_launch_ms": g_ms / g_cnt, "share_of_sweep": g_ms / tot if tot else None,
                               "kernel_ms_per_sweep": {nm: v[1] / max(v[0], 1) for nm, v in prof.items()}}
    return out


def roll_cpu_leg(n_windows=8, chains=4, sweeps=2, iters_ref=200):
    """The CPU path beside it: oracle, reference-faithful (BLAS) and algorithm-matched formulations, (window, chain) units in
    the collapse(2) shape of forecaster.h:626 (one OpenMP task per unit).  windows/s = unit-sweeps/s / (chains x sweeps per refit)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import binding as ob
    m, pk, y = roll_problem(0, n_windows, chains, 4, 0)
    thr = host_threads()
    res = {"cores": thr, "sample": f"{n_windows} windows x {chains} chains of C4, {sweeps} timed sweeps after 1 warm-up; windows/s at {iters_ref} sweeps per refit"}
    for name, faithful, blas in (("faithful", True, True), ("efficient", False, False)):
        backend = ob.use_scipy_openblas(blas)
        f = ob.OracleFit(pk, faithful=faithful, threads=thr)
        f.step(1, False)
        t0 = time.perf_counter(); f.step(sweeps, False); dt = time.perf_counter() - t0
        f.close()
        us = n_windows * chains * sweeps / dt
        res[name] = {"unit_sweeps_per_s": us, "windows_per_s": us / (chains * iters_ref), "dense_kernels": backend}
    ob.use_scipy_openblas(False)
    return res


# ------------------------------------------------------------------------------------------------ the CPU arm
def cpu_arm(mode_faithful: bool, chains: int, sweeps: int, warm: int, threads: int = 0, blas: bool = True):
    """Time the CPU oracle on a bounded sample of the C3 workload; returns chain-sweeps/s.  Parallel shape = the
    reference's: one OpenMP task per chain (triangular.h:1222), chains >= threads."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import binding as ob  # test infrastructure; only the baseline legs of bench.py may use it
    m = build_model(chains, 10, 0, seed=7, device=0)
    pk = m._pack(m.design_, m.response_, m._seeds(1, chains))
    threads = threads or host_threads()  # (torchrun exports OMP_NUM_THREADS=1: do not rely on it)
    backend = ob.use_scipy_openblas(blas and mode_faithful)
    f = ob.OracleFit(pk, faithful=mode_faithful, threads=threads)
    if warm:
        f.step(warm, False)
    t0 = time.perf_counter()
    f.step(sweeps, True)  # recorded sweeps, like the GPU arm's timed region
    dt = time.perf_counter() - t0
    f.close()
    ob.use_scipy_openblas(False)
    return chains * sweeps / dt, threads, dt, backend


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    thr = host_threads()
    chains = max(thr, 8)  # one chain per host thread: the reference's parallel shape
    v, cores, dt, backend = cpu_arm(True, chains, args.steps, args.warmup)
    veff, _, _, _ = cpu_arm(False, 4 * chains, 3, 1)
    sample = (f"{chains} chains (one OpenMP task per chain, triangular.h:1222) x {args.steps} recorded sweeps after {args.warmup} warm-up, "
              f"reference-faithful formulation (Kronecker design tri.h:286, dense n x n SV precision draw.h:454-487); dense kernels: {backend}")
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.chains, args.prior, world),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, "sample_chains": chains,
                         "value_efficient_formulation": veff,
                         "sample_efficient": f"{4 * chains} chains x 3 sweeps, weighted-Gram + tridiagonal formulation (the algorithm the GPU runs)"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ the GPU arm
def timed_sweeps(fit, K, W, barrier, record=True):
    fit.step(W, False)  # warm-up: first sweep eager, then the captured graph
    if record:
        fit.step(1, True)  # (the recording graph is captured outside the timed region too)
    barrier()
    l0 = fit.launch_count()
    fit.step(K, record, sync=False)
    fit.sync()
    barrier()
    return fit.last_ms(), fit.launch_count() - l0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--chains", type=int, default=CHAINS_PER_GPU, help="chains per GPU")
    ap.add_argument("--prior", default="DL")
    ap.add_argument("--roll-windows", type=int, default=32, help="rolling windows per GPU of the forecast_roll metric")
    ap.add_argument("--roll-iters", type=int, default=200, help="sweeps per refit of the forecast_roll metric (half of them burn-in)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-roll", action="store_true")
    ap.add_argument("--no-strong", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = max(args.warmup, 1) if args.impl == "reference" else 3
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from bvhar_b200._abi import REC_COEF, REC_LVOL_LAST
    from bvhar_b200._engine import CudaFit

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the bvhar_b200 engine has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    K, W = args.steps, args.warmup
    REC = REC_COEF | REC_LVOL_LAST  # what the forecasters and summaries read; the 164 KB / draw / chain h_record is opt-in
    model = build_model(args.chains, W + 2 * K + 8, 0, seed=1000 + rank, device=local, prior=args.prior)
    k, d, n = model.n_features_in_, model.design_.shape[1], model.response_.shape[0]
    packed = model._pack(model.design_, model.response_, model._seeds(1, args.chains), record_mask=REC, unit_id_base=rank * args.chains)
    fit = CudaFit(packed)
    sampler = ClockSampler(local)
    sampler.start()     # nvidia-smi is looping before the timed region starts (its start-up can exceed a 150 ms region)
    n0 = len(sampler.rows)
    ms, launches = timed_sweeps(fit, K, W, barrier)
    ms = allmax(ms)
    n1 = len(sampler.rows)
    # the timed region may be shorter than the 100 ms sampling period: the same load (untimed sweeps) is kept up until at
    # least three samples have been taken under it
    t_lim = time.perf_counter() + 3.0
    while len(sampler.rows) - n1 < 3 and time.perf_counter() < t_lim:
        fit.step(K, False)
    sampler.rows = sampler.rows[n0:]
    clocks = sampler.stop()
    clocks["samples_in_timed_region"] = n1 - n0
    # same K steps again with a CUDA-event pair around every kernel (eager launches) for the roofline
    fit.set_profiling(True)
    fit.step(K, False)
    prof = fit.profile()
    fit.set_profiling(False)
    fit.close()
    value = world * args.chains * K / (ms * 1e-3)

    # ---- strong scaling: 1,024 chains in total, split over the ranks (global Philox keys) ------------------
    strong = None
    if not args.no_strong:
        cs = STRONG_CHAINS_TOTAL // world
        if world == 1 and cs == args.chains:
            strong = {"chains_total": STRONG_CHAINS_TOTAL, "chains_per_gpu": cs, "value": value, "unit": UNIT, "ms_per_step": ms / K}
        else:
            ms_s = build_model(cs, W + K + 4, 0, seed=1500 + rank, device=local, prior=args.prior)
            pk_s = ms_s._pack(ms_s.design_, ms_s.response_, ms_s._seeds(1, cs), record_mask=REC, unit_id_base=rank * cs)
            fs = CudaFit(pk_s)
            t_s, _ = timed_sweeps(fs, K, W, barrier)
            fs.close()
            t_s = allmax(t_s)
            strong = {"chains_total": cs * world, "chains_per_gpu": cs, "value": world * cs * K / (t_s * 1e-3), "unit": UNIT,
                      "ms_per_step": t_s / K}

    # ---- end to end through the public API: host buffers in, draws out -----------------------
    e2e = None
    if not args.no_e2e:
        # one untimed call first: CUDA module load and the page-locked output pool (bvhar_b200/_hostpool.py) are
        # set up by the first fit of a process; its results are dropped so the pooled buffers are free again
        import gc
        m_warm = build_model(args.chains, W + K, W, seed=3000 + rank, device=local, prior=args.prior)
        m_warm.fit()
        del m_warm
        gc.collect()
        dts = []
        for rep in range(3):  # three timed fits, each a fresh model and fresh outputs; the median is reported
            m2 = build_model(args.chains, W + K, W, seed=2000 + 10 * rep + rank, device=local, prior=args.prior)
            h2d = m2.design_.nbytes + m2.response_.nbytes + sum(sum(np.asarray(v).nbytes for v in ini.values()) for ini in m2.init_)
            barrier()
            t0 = time.perf_counter()
            m2.fit()
            torch.cuda.synchronize()
            dts.append(allmax(time.perf_counter() - t0))
            d2h = sum(a.nbytes for a in m2.param_.values())
            rec_names = sorted(m2.param_.keys())
            del m2
            gc.collect()
        dt = sorted(dts)[1]
        e2e = {"value": world * args.chains * (W + K) / dt, "unit": UNIT, "h2d_bytes_per_step": int(h2d / (W + K)),
               "d2h_bytes_per_step": int(d2h / K), "sweeps": W + K, "recorded_draws": K, "seconds": dt, "seconds_all": dts,
               "records": rec_names,
               "note": "VharBayes.fit(), default record policy (every record; for >= 64 SV chains the last-time-point h instead of the "
                       "full n x k h_record, which stays opt-in): pack + H2D + (warm-up + K) sweeps + D2H of every record into pooled "
                       "page-locked host memory, streamed out while later sweeps sample (pool warmed by one untimed fit; median of 3)"}

    # ---- second metric of BASELINE.json: forecast_roll windows/s (C4) --------------
    roll = None
    if not args.no_roll:
        try:
            roll = roll_metric(local, rank, world, args.roll_windows, 4, args.roll_iters, args.roll_iters // 2, barrier, allmax)
        except Exception as e:  # the headline line must still be printed
            roll = {"error": str(e)[:300]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    fl = algorithmic_flops(k, d, n)
    peaks, peak_src = measured_peaks()
    gname = "gram_S_dmma"
    g_cnt, g_ms = prof.get(gname, (0, 0.0))
    g_avg_ms = g_ms / max(g_cnt, 1)
    tot_prof_ms = sum(v[1] for v in prof.values())
    achieved = args.chains * fl["gram"] / (g_avg_ms * 1e-3) / 1e12 if g_avg_ms > 0 else None
    f64 = fp64_peaks(local)
    cublas_peak = f64["cublas_dgemm_tflops"]
    roofline = {
        "kernel": "dgemm_tn_kernel (G1: S = Dinv' XX, SV-weighted Gram of every (chain, equation))",
        "bound": "tensor", "achieved": achieved, "peak": cublas_peak, "unit": "TFLOP/s",
        "frac": achieved / cublas_peak if achieved else None, "traffic": ncu_traffic(),
        "peak_source": "fp64 is absent from MEASURED_PEAKS.json: cuBLAS DGEMM 8192^3 measured in this run (burst, best of 3); "
                       f"DMMA issue peak {f64['dmma_issue_tflops']:.1f} TFLOP/s, DFMA peak {f64['dfma_tflops']:.1f} TFLOP/s (bench_kernels.cu); "
                       f"HBM {peaks.get('hbm_gbs')} GB/s ({peak_src})",
        "fp64_peaks": f64,
        "algorithmic_flops_per_launch": args.chains * fl["gram"], "avg_launch_ms": g_avg_ms,
        "share_of_step": g_ms / tot_prof_ms if tot_prof_ms > 0 else None,
        "sweep_frac_of_peak": (args.chains * fl["total"] / (ms / K * 1e-3) / 1e12) / cublas_peak,
        "kernel_ms_per_step": {nm: v[1] / K for nm, v in prof.items()},
    }
    if roll and "roofline" in roll:
        roll["roofline"]["peak"] = cublas_peak
        roll["roofline"]["frac"] = roll["roofline"]["achieved"] / cublas_peak
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        # in a fresh process (this one carries torch's / CUDA's thread pools), exactly as the --impl reference arm runs it
        env = {k2: v2 for k2, v2 in os.environ.items() if k2 not in ("OMP_NUM_THREADS", "RANK", "WORLD_SIZE", "LOCAL_RANK")}
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "3", "--warmup", "1"],
                             capture_output=True, text=True, env=env, timeout=900).stdout.strip().splitlines()[-1]
        cpu = json.loads(out)["cpu_baseline"]
        if roll is not None and "error" not in roll:
            code = ("import json, bench; print(json.dumps(bench.roll_cpu_leg(iters_ref=%d)))" % args.roll_iters)
            o2 = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=900, cwd=ROOT)
            try:
                roll["cpu_baseline"] = json.loads(o2.stdout.strip().splitlines()[-1])
            except Exception:
                roll["cpu_baseline"] = {"error": (o2.stderr or o2.stdout)[-300:]}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.chains, args.prior, world),
        "algorithmic_gflop_per_step": args.chains * fl["total"] / 1e9,
        "sustained_tflops_algorithmic": world * args.chains * fl["total"] * K / (ms * 1e-3) / 1e12,
        "clocks": clocks, "gpu_launches": int(launches), "e2e": e2e, "strong": strong, "roofline": roofline, "cpu_baseline": cpu,
        "forecast_roll": roll,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py — Gibbs sweeps/s of the B200 engine on BASELINE.json's headline configuration.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[2], "C3"): synthetic 20-variable VHAR(5,22) with stochastic
volatility, Dirichlet-Laplace prior, 1,024 chains per GPU (k=20, d=61, n=1000).  One *step* = one
Gibbs sweep of every chain.  `value` = chain-sweeps per second over all GPUs, timed on the device
(CUDA events on the engine's stream, max over ranks) with all inputs resident in HBM; `e2e` = the same
metric through the public API (`VharBayes.fit()`): host buffers in, posterior draws out, copies inside
the timed region.  The `--impl reference` arm times the CPU restatement of the reference's own
algorithm (oracle/, faithful mode: Kronecker design + dense n x n SV precision) on all host cores.
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


def build_model(chains, iters, burn, seed, device, record_mask=None, prior="DL", dim=20, T=1022):
    from bvhar_b200 import _spec as sp
    from bvhar_b200 import datasets
    from bvhar_b200._abi import REC_ALL
    from bvhar_b200.model import VharBayes
    cfg = {"DL": sp.DlConfig, "Horseshoe": sp.HorseshoeConfig, "NG": sp.NgConfig, "SSVS": sp.SsvsConfig,
           "GDP": sp.GdpConfig, "Minnesota": lambda: sp.MinnesotaConfig(is_long=True)}[prior]()
    y = datasets.simulate_vhar_sv(T=T, dim=dim, seed=20240001)
    return VharBayes(y, week=5, month=22, n_chain=chains, n_iter=iters, n_burn=burn, bayes_config=cfg,
                     cov_config=sp.SvConfig(), seed=seed, device=device,
                     record_mask=REC_ALL if record_mask is None else record_mask)


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
    """dram__bytes_read.sum + dram__bytes_write.sum of the G1 launch from the committed `ncu --set full` capture
    (profiles/r01_gram_traffic.json: same workload, 1 024 chains), or None."""
    p = os.path.join(ROOT, "profiles", "r01_gram_traffic.json")
    try:
        with open(p) as f:
            return json.load(f)["traffic_bytes_per_launch"]
    except Exception:
        return None


def roll_metric(device, n_windows=256, chains=4, iters=12, burn=4):
    """forecast_roll() on BASELINE config C4 (synthetic 50-variable VAR(4), NG prior + SV, 4 chains per window), a bounded
    sample of the 2 000 windows: every (window, chain) refitted and forecast one step ahead through the native
    out-of-sample runner, host buffers in, forecast draws out.  windows/s scales as 1 / iters, so the sweep rate is
    reported too, and windows/s is extrapolated to the 1 000-iteration setting of SURVEY.md section 8(d)."""
    from bvhar_b200 import _spec as sp
    from bvhar_b200 import datasets
    from bvhar_b200._engine import outforecast
    from bvhar_b200.model import VarBayes
    y = datasets.simulate_var_sv(404 + n_windows, 50, 4)
    m = VarBayes(y[:404], 4, chains, iters, burn, 1, sp.NgConfig(), sp.SvConfig(), seed=1, device=device)
    x_tot, y_tot = m._design_of(y)
    n0 = m.response_.shape[0]
    pk = m._pack(x_tot, y_tot, m._seeds(n_windows, chains), win_row0=list(range(1, n_windows + 1)), win_nrow=[n0] * n_windows,
                 record_mask=0, unit_id_base=chains)
    outforecast(pk, y, 1, 4, m._seeds(chains))  # warm-up (module load, memory pools)
    t0 = time.perf_counter()
    fc, _ = outforecast(pk, y, 1, 4, m._seeds(chains))
    dt = time.perf_counter() - t0
    assert np.all(np.isfinite(fc))
    return {"metric": "forecast_roll_windows_per_s", "config": f"C4 sample: {n_windows} rolling windows x {chains} chains, k=50 d=201 n=400, "
            f"NG prior + SV, {iters} sweeps per refit ({burn} burn-in), 1-step forecast", "value": n_windows / dt, "unit": "windows/s",
            "seconds": dt, "unit_sweeps_per_s": n_windows * chains * iters / dt,
            "windows_per_s_at_1000_iters": n_windows * iters / dt / 1000.0, "path": "bvhar_cuda_outforecast_run (end to end, host buffers)"}


def cpu_arm(mode_faithful: bool, chains: int, sweeps: int, warm: int, threads: int = 0):
    """Time the CPU oracle on a bounded sample of the same workload; returns chain-sweeps/s."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import binding as ob  # test infrastructure; only this baseline leg of bench.py may use it
    m = build_model(chains, 10, 0, seed=7, device=0)
    pk = m._pack(m.design_, m.response_, m._seeds(1, chains))
    if threads == 0:  # every host core this process may use (torchrun exports OMP_NUM_THREADS=1: do not rely on it)
        try:
            threads = len(os.sched_getaffinity(0))
        except AttributeError:
            threads = os.cpu_count() or 1
    f = ob.OracleFit(pk, faithful=mode_faithful, threads=threads)
    if warm:
        f.step(warm, False)
    t0 = time.perf_counter()
    f.step(sweeps, False)
    dt = time.perf_counter() - t0
    cores = threads
    f.close()
    return chains * sweeps / dt, cores, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ncpu = os.cpu_count() or 1
    chains = 2  # bounded sample: 2 chains swept with every host thread (series-level inner parallelism)
    v, cores, dt = cpu_arm(True, chains, args.steps, args.warmup)
    veff, _, _ = cpu_arm(False, max(ncpu, 8), 3, 1)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C3: synthetic 20-variable VHAR(5,22)-SV, DL prior, k=20 d=61 n=1000", "sample_chains": chains},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{chains} chains x {args.steps} sweeps after {args.warmup} warm-up, reference-faithful formulation "
                                   "(Kronecker design tri.h:286, dense n x n SV precision draw.h:454-487), hand-written loops, no BLAS",
                         "value_efficient_formulation": veff},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--chains", type=int, default=1024, help="chains per GPU")
    ap.add_argument("--prior", default="DL")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-roll", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = max(args.warmup, 1) if args.impl == "reference" else 3
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from bvhar_b200._abi import REC_ALL
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

    K, W = args.steps, args.warmup
    model = build_model(args.chains, W + 2 * K + 8, 0, seed=1000 + rank, device=local, prior=args.prior)
    k, d, n = model.n_features_in_, model.design_.shape[1], model.response_.shape[0]
    packed = model._pack(model.design_, model.response_, model._seeds(1, args.chains), record_mask=0)
    fit = CudaFit(packed)
    sampler = ClockSampler(local)
    sampler.start()     # nvidia-smi is looping before the timed region starts (its start-up can exceed a 150 ms region)
    fit.step(W, False)  # warm-up: first sweep eager, then the captured graph
    barrier()
    n0 = len(sampler.rows)
    l0 = fit.launch_count()
    fit.step(K, False, sync=False)
    fit.sync()
    barrier()
    ms = fit.last_ms()
    launches = fit.launch_count() - l0
    n1 = len(sampler.rows)
    # the timed region may be shorter than the 100 ms sampling period: the same load (untimed sweeps) is kept up until at
    # least three samples have been taken under it
    t_lim = time.perf_counter() + 3.0
    while len(sampler.rows) - n0 < 3 and time.perf_counter() < t_lim:
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
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * args.chains * K / (ms * 1e-3)

    # ---- end to end through the public API: host buffers in, draws out -----------------------
    e2e = None
    if not args.no_e2e:
        # one untimed call first: CUDA module load and the page-locked output pool (bvhar_b200/_hostpool.py) are
        # set up by the first fit of a process; its results are dropped so the pooled buffers are free again
        import gc
        m_warm = build_model(args.chains, W + K, W, seed=3000 + rank, device=local, prior=args.prior, record_mask=REC_ALL)
        m_warm.fit()
        del m_warm
        gc.collect()
        dts = []
        for rep in range(3):  # three timed fits, each a fresh model and fresh outputs; the median is reported
            m2 = build_model(args.chains, W + K, W, seed=2000 + 10 * rep + rank, device=local, prior=args.prior, record_mask=REC_ALL)
            h2d = m2.design_.nbytes + m2.response_.nbytes + sum(sum(np.asarray(v).nbytes for v in ini.values()) for ini in m2.init_)
            barrier()
            t0 = time.perf_counter()
            m2.fit()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            d2h = sum(a.nbytes for a in m2.param_.values())
            te = torch.tensor([dt], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(te, op=dist.ReduceOp.MAX)
            dts.append(float(te.item()))
            del m2
            gc.collect()
        dt = sorted(dts)[1]
        e2e = {"value": world * args.chains * (W + K) / dt, "unit": UNIT, "h2d_bytes_per_step": int(h2d / (W + K)),
               "d2h_bytes_per_step": int(d2h / K), "sweeps": W + K, "recorded_draws": K, "seconds": dt, "seconds_all": dts,
               "note": "VharBayes.fit(): pack + H2D + (warm-up + K) sweeps + D2H of every record incl. h_record into pooled page-locked "
                       "host memory, streamed out while later sweeps sample (pool warmed by one untimed fit; median of 3 timed fits)"}

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
    # fp64 peak: not in MEASURED_PEAKS.json -> measure the DMMA issue peak and cuBLAS DGEMM here
    import ctypes as C
    from bvhar_b200 import _abi
    L = _abi.load_library()
    pk = C.c_double()
    L.bvhar_cuda_bench_fp64_peak(local, 1, C.byref(pk))
    dmma_peak = pk.value
    a = torch.randn(8192, 8192, dtype=torch.float64, device="cuda"); b = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
    torch.matmul(a, b); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for _ in range(3):
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    cublas_peak = 2 * 8192 ** 3 / (best * 1e-3) / 1e12
    del a, b
    roofline = {
        "kernel": "dgemm_tn_kernel (G1: S = Dinv' XX, SV-weighted Gram of every (chain, equation))",
        "bound": "tensor", "achieved": achieved, "peak": cublas_peak, "unit": "TFLOP/s",
        "frac": achieved / cublas_peak if achieved else None, "traffic": ncu_traffic(),
        "peak_source": "fp64 is absent from MEASURED_PEAKS.json: cuBLAS DGEMM 8192^3 measured in this run (burst, best of 3); "
                       f"DMMA issue peak {dmma_peak:.1f} TFLOP/s; HBM {peaks.get('hbm_gbs')} GB/s ({peak_src})",
        "algorithmic_flops_per_launch": args.chains * fl["gram"], "avg_launch_ms": g_avg_ms,
        "share_of_step": g_ms / tot_prof_ms if tot_prof_ms > 0 else None,
        "kernel_ms_per_step": {nm: v[1] / K for nm, v in prof.items()},
    }
    # ---- second metric of BASELINE.json: forecast_roll windows/s (C4, bounded sample) --------------
    roll = None
    if world == 1 and not args.no_roll:
        try:
            roll = roll_metric(local)
        except Exception as e:  # the headline line must still be printed
            roll = {"error": str(e)[:200]}
    cpu = None
    if not args.no_cpu_baseline:
        ncpu = os.cpu_count() or 1
        # in a fresh process (this one carries torch's / CUDA's thread pools), exactly as the --impl reference arm runs it
        env = {k2: v2 for k2, v2 in os.environ.items() if k2 not in ("OMP_NUM_THREADS", "RANK", "WORLD_SIZE", "LOCAL_RANK")}
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "3", "--warmup", "1"],
                             capture_output=True, text=True, env=env, timeout=900).stdout.strip().splitlines()[-1]
        ref = json.loads(out)
        vf, cores, ve = ref["value"], ref["cpu_baseline"]["cores"], ref["cpu_baseline"]["value_efficient_formulation"]
        cpu = {"value": vf, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "2 chains x 3 sweeps after 1 warm-up of the same workload, reference-faithful formulation "
                         "(Kronecker design, dense n x n SV precision), all host threads; hand-written loops, no BLAS",
               "value_efficient_formulation": ve,
               "sample_efficient": f"{max(ncpu, 8)} chains x 3 sweeps, weighted-Gram + tridiagonal formulation"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"C3: synthetic 20-variable VHAR(5,22)-SV, {args.prior} prior, {args.chains} chains per GPU, k={k} d={d} n={n}",
                   "chains_total": world * args.chains, "l2": "per-sweep working set (>1 GB of chain state) exceeds the 126 MB L2",
                   "timed_path": "CUDA graph replay of the sweep; roofline from a second pass with per-kernel events"},
        "algorithmic_gflop_per_step": args.chains * fl["total"] / 1e9,
        "sustained_tflops_algorithmic": world * args.chains * fl["total"] * K / (ms * 1e-3) / 1e12,
        "clocks": clocks, "gpu_launches": int(launches), "e2e": e2e, "roofline": roofline, "cpu_baseline": cpu, "forecast_roll": roll,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

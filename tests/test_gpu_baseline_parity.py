"""GPU engine vs CPU oracle at the EXACT shapes of the five BASELINE.json configurations, natural kernel dispatch
(no BVHAR_COEF_BIG override): the kernels the bench times - `dgemm_tn_kernel` full-tile path at M = chains x 20,
`coef_pbuild_kernel<3,6>`, `coef_chol_kernel_nc2`, `coef_seq_pipe_kernel<2>`, `impact_sv_kernel`, `sv_*<24>` (C3);
`coef_chol_cta_kernel`, `coef_seq_time_kernel`, `coef_pbuild_kernel<4,16>` (C4: k=50, d=201); the k=100 / d=301 LDLT
path `coef_seq_big_kernel` (C5) - are compared state for state with the oracle's efficient formulation (itself
proved equal to the reference-faithful one in tests/test_oracle_pins.py and pinned by tests/test_numpy_mirror.py).

Tolerances (north_star level 1): every stage of sweeps 1 and 2 is run on both sides FROM IDENTICAL STATES (after each
stage the oracle's outputs are written into the engine, `bvhar_cuda_fit_set_state`, so no stage inherits the rounding of
the stages before it) and must agree to 1e-10 relative, with two documented, COMPUTED exceptions:

  * the coefficient draw (stage 4) solves P_j a = b.  Two backward-stable fp64 Cholesky solvers (the device's blocked
    one, the host's, Eigen's) agree only to ~ d * cond_s(P_j) * 2^-53, cond_s = the condition number of the diagonally
    equilibrated matrix (van der Sluis: the quantity Cholesky's forward error depends on).  The real-data configurations
    are collinear (etf_vix levels: cond_s = 2.5e5 at C1, 1.6e5 at C2; synthetic C3 / C4: 60 / 1e3), so stage 4 is held
    to max(1e-10, 16 * d * cond_s * 2^-53), cond_s computed here by NumPy from an independent reconstruction of
    P_j = diag(prec_j) + X' diag(sum_i L_ij^2 / s_ti^2) X.
  * the inverse-Gaussian draws (GDP local scales, DL latents) evaluate the reference's root formula literally
    (common.h:104-106), which cancels catastrophically when mean * z^2 >> shape: its rounding error is
    ~ 2^-53 * (mean / root)^2 (the small root is mean * (1 + (w - r) / (2 shape)) with w + r = 2 shape (mean / root - 1)),
    different for any two fp64 evaluations (see oracle/oracle_rng.h why the literal form is kept).  Those elements are
    held to max(1e-10, 4 * amp_i^2 * 2^-53), amp_i = max(mean_i / x_i, x_i / mean_i) computed here from the oracle's states.

Whole records after 3 recorded sweeps (CUDA-graph path, nothing re-synchronised): 1e-7, scaled by the stage-4 factor.
The oracle sweeps these sizes in well under a second per unit.
"""
import numpy as np
import pytest

import baseline_configs as bc
import oracle_binding as ob
from bvhar_b200 import _spec as sp
from bvhar_b200._engine import CudaFit
from test_gpu_parity import HYPER, STATES, relerr

pytestmark = pytest.mark.gpu

STAGE_TOL = 1e-10
TRAJ_TOL = 1e-7
PRIOR_CFG = {"SSVS": sp.SsvsConfig, "Horseshoe": sp.HorseshoeConfig, "DL": sp.DlConfig, "NG": sp.NgConfig, "GDP": sp.GdpConfig}

# (configuration, prior, chains, windows): the prior BASELINE.json names for the configuration first, then the other
# shrinkage priors at the headline shape C3
CASES = [("C1", "SSVS", 2, 1), ("C2", "Horseshoe", 2, 1), ("C3", "DL", 2, 1), ("C4", "NG", 2, 3), ("C5", "GDP", 1, 1),
         ("C3", "Horseshoe", 2, 1), ("C3", "SSVS", 2, 1), ("C3", "NG", 2, 1), ("C3", "GDP", 2, 1)]


def _names(stage, prior, sv):
    names = list(STATES.get(stage, []))
    if stage in (1, 5):
        names += HYPER[prior]
    if stage == 9:
        names += ["lvol_draw", "lvol_sig", "lvol_init"] if sv else ["diag_vec"]
    if stage == 3 and not sv:
        names = []
    return names


# states written back into the engine after a stage (everything later stages read; the SAVS copies and sqrt_sv are
# derived, never read back)
def _sync_names(stage, prior, sv):
    return [nm for nm in _names(stage, prior, sv) if nm not in ("sparse_coef", "sparse_contem", "sqrt_sv")]


def coef_precision_cond(pk, o, w, c):
    """max_j cond_2 of the diagonally equilibrated P_j of unit (w, c), P_j rebuilt by NumPy from the oracle's current
    sqrt_sv / chol_lower / prior_alpha_prec (tri.h:281-316 in weighted-Gram form)."""
    k, d = pk.dim, pk.dim_design
    n, r0 = int(pk.win_nrow[w]), int(pk.win_row0[w])
    X = np.asarray(pk.X)[r0:r0 + n]
    s = o.get_state("sqrt_sv", w, c).reshape(n, k, order="F")
    L = o.get_state("chol_lower", w, c).reshape(k, k, order="F")
    pr = o.get_state("prior_alpha_prec", w, c)
    N = pk.n_alpha
    nrow = N // k
    worst = 1.0
    for j in range(k):
        wt = ((L[j:, j] ** 2)[None, :] / s[:, j:] ** 2).sum(axis=1)
        P = X.T @ (X * wt[:, None])
        pj = np.concatenate([pr[j * nrow:(j + 1) * nrow], [pr[N + j]]]) if pk.include_mean else pr[j * d:(j + 1) * d]
        P += np.diag(pj)
        dg = 1.0 / np.sqrt(np.diag(P))
        ev = np.linalg.eigvalsh(P * dg[:, None] * dg[None, :])
        worst = max(worst, float(ev[-1] / ev[0]))
    return worst


def stage4_tol(cond, d):
    return max(STAGE_TOL, 16.0 * d * cond * 2.0 ** -53)


def invgauss_amp(prior, stage, name, pk, o, w, c):
    """Per-element max(mean / x, x / mean) of the inverse-Gaussian variate x behind state `name` (the states hold 1 / x; the
    rounding error of the literal root formula, common.h:104-106, is ~ 2^-53 times its square), or None when the state
    does not come from an inverse-Gaussian draw."""
    N = pk.n_alpha
    if prior == "GDP" and stage == 1 and name in ("local_lev", "prior_alpha_prec"):
        # local_i = 1 / InvGauss(|rho_g(i) / alpha_i|, rho^2), prec = 1 / local (tri.h:985, draw.h:1147-1153)
        local = o.get_state("local_lev", w, c)
        prec = o.get_state("prior_alpha_prec", w, c)
        alpha = o.get_state("coef_mat", w, c).reshape(pk.dim_design, pk.dim, order="F")[: N // pk.dim].reshape(-1, order="F")
        # rho_g(i) is not exposed per coefficient; the mean follows from the two roots: x1 * x2 = mean^2 and the chosen
        # root x = 1 / local: amp = max(mean / x, x / mean) >= 1 either way, bounded through |alpha| and the group rates
        rates = o.get_state("group_rate", w, c)
        amp = np.maximum.reduce([np.maximum(r / np.abs(alpha) * local, 1.0 / (r / np.abs(alpha) * local)) for r in rates])
        return np.concatenate([amp, np.ones(prec.size - N)]) if name == "prior_alpha_prec" else amp
    if prior == "GDP" and stage == 5 and name in ("contem_fac", "prior_chol_prec"):
        fac, rate, a = o.get_state("contem_fac", w, c), o.get_state("contem_rate", w, c), o.get_state("contem_coef", w, c)
        m = np.abs(rate / a)
        return np.maximum(m * fac, 1.0 / (m * fac))
    if prior == "DL" and stage == 1 and name in ("latent_local", "prior_alpha_prec"):
        # latent_i = 1 / InvGauss(sc_i / |alpha_i|, 1), prec_i = 1 / (sc_i^2 latent_i)  (tri.h:899-900, draw.h:758-770)
        lat = o.get_state("latent_local", w, c)
        prec = o.get_state("prior_alpha_prec", w, c)
        alpha = o.get_state("coef_mat", w, c).reshape(pk.dim_design, pk.dim, order="F")[: N // pk.dim].reshape(-1, order="F")
        sc = 1.0 / np.sqrt(prec[:N] * lat)
        m = sc / np.abs(alpha)
        amp = np.maximum(m * lat, 1.0 / (m * lat))
        return np.concatenate([amp, np.ones(prec.size - N)]) if name == "prior_alpha_prec" else amp
    if prior == "DL" and stage == 5 and name in ("latent_contem_local", "prior_chol_prec"):
        lat, loc, a = o.get_state("latent_contem_local", w, c), o.get_state("contem_local_lev", w, c), o.get_state("contem_coef", w, c)
        m = loc / np.abs(a)
        return np.maximum(m * lat, 1.0 / (m * lat))
    return None


def elem_err(a, b, tol):
    """max over elements of |a - b| / (tol_i * max(|a_i|, |b_i|)): <= 1 means every element is within its own tolerance."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape and np.array_equal(np.isnan(a), np.isnan(b))
    ok = np.isfinite(a) & np.isfinite(b)
    sc = np.maximum(np.abs(a[ok]), np.abs(b[ok])) + 1e-300
    return float(np.max(np.abs(a[ok] - b[ok]) / (np.broadcast_to(tol, a.shape)[ok] * sc))) if ok.any() else 0.0


@pytest.mark.parametrize("cfg,prior,chains,windows", CASES, ids=[f"{c}-{p}" for c, p, _, _ in CASES])
def test_stage_parity_at_baseline_shape(cfg, prior, chains, windows):
    """Every stage of sweeps 1 and 2, each from states identical on both sides, every unit."""
    pk = bc.packed(cfg, chains=chains, windows=windows, iters=3, prior_cfg=PRIOR_CFG[prior]())
    assert (pk.dim, pk.dim_design, int(pk.win_nrow[0])) == bc.SHAPES[cfg]
    sv = bc.SV[cfg]
    o = ob.OracleFit(pk, faithful=False); g = CudaFit(pk)
    units = [(w, c) for w in range(pk.n_windows) for c in range(pk.n_chains)]
    worst, conds = {}, []
    for it in (1, 2):
        for st in range(1, 10):
            tol = STAGE_TOL
            if st == 4:
                cond = max(coef_precision_cond(pk, o, w, c) for w, c in units)
                tol = stage4_tol(cond, pk.dim_design)
                conds.append(cond)
            o.run_stage(st, it); g.run_stage(st, it)
            for nm in _names(st, prior, sv):
                for w, c in units:
                    amp = invgauss_amp(prior, st, nm, pk, o, w, c)
                    if amp is not None:  # element-wise: each inverse-Gaussian variate against its own amplification
                        e = elem_err(o.get_state(nm, w, c), g.get_state(nm, w, c), np.maximum(STAGE_TOL, 4.0 * amp * amp * 2.0 ** -53))
                        assert e <= 1.0, f"{cfg} {prior} sweep {it} stage {st} {nm}[w{w} c{c}]: {e:.2f} x its inverse-Gaussian tolerance"
                        continue
                    e = relerr(o.get_state(nm, w, c), g.get_state(nm, w, c))
                    worst[(it, st, nm)] = max(worst.get((it, st, nm), 0.0), e)
                    assert e < tol, f"{cfg} {prior} sweep {it} stage {st} {nm}[w{w} c{c}]: rel err {e:.3e} (tol {tol:.1e})"
            for nm in _sync_names(st, prior, sv):
                for w, c in units:
                    g.set_state(nm, o.get_state(nm, w, c), w, c)
    top = max(worst, key=worst.get)
    print(f"{cfg} {prior}: cond_s(P) {max(conds):.1e}, stage-4 tol {stage4_tol(max(conds), pk.dim_design):.1e}; worst error {worst[top]:.2e} at {top}")
    o.close(); g.close()


@pytest.mark.parametrize("cfg,prior,chains,windows", CASES[:5], ids=[f"{c}-{p}" for c, p, _, _ in CASES[:5]])
def test_three_recorded_sweeps_at_baseline_shape(cfg, prior, chains, windows):
    """Three recorded sweeps through the CUDA-graph path: every record of every unit within 1e-7 of the oracle's
    (times the conditioning factor of the coefficient solves, see the module docstring)."""
    pk = bc.packed(cfg, chains=chains, windows=windows, iters=3, burn=0, prior_cfg=PRIOR_CFG[prior]())
    o = ob.OracleFit(pk, faithful=False); g = CudaFit(pk)
    o.run_stage(1, 1); o.run_stage(3, 1)
    cond = max(coef_precision_cond(pk, o, w, c) for w in range(pk.n_windows) for c in range(pk.n_chains))
    o.close()
    tol = TRAJ_TOL * stage4_tol(cond, pk.dim_design) / STAGE_TOL
    o = ob.OracleFit(pk, faithful=False)
    o.run(); g.run()
    assert g.record_names() == o.record_names()
    for nm in g.record_names():
        for w in range(pk.n_windows):
            for c in range(pk.n_chains):
                a, b = o.record(nm, w, c), g.record(nm, w, c)
                assert a.shape == b.shape and a.shape[0] == 3
                e = relerr(a, b)
                assert e < tol, f"{cfg} {prior} {nm}[w{w} c{c}]: rel err {e:.3e} (tol {tol:.1e})"
    o.close(); g.close()

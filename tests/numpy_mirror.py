"""Independent NumPy restatement of ONE Gibbs sweep of the reference sampler, written from the reference's text only
(`/root/reference/inst/include/bvhar/src/bayes/triangular/triangular.h:40-71, 102-115, 254-262, 281-348, 371-372,
400-409, 599-623, 885-908` and `bayes/misc/draw.h:99-132, 269-361, 372-430, 440-537, 758-890, 1244-1259`,
`triangular/config.h:62-104, 243-248, 327, 371-420, 457-475`), in the reference's own DENSE formulation: Kronecker design per
equation, dense n x n stochastic-volatility precision, `numpy.linalg` Cholesky / solves.

TEST INFRASTRUCTURE.  Purpose (VERDICT r1 "common-mode risk"): the CUDA engine and the C++ oracle share the Python
packing (`_abi.PackedFit`) and were written by the same hand from the same restatement.  This mirror shares neither: it
reads the *unpacked* keyword arguments (the `LIST`s of the reference's constructors) and re-derives every conditional
from the reference.  Only the random variates are injected: each draw site asks `oracle_draw_at` for the variate at its
Philox address (the stream is a specification of this repo, not of the reference; the samplers themselves are pinned
against SciPy in tests/test_oracle_samplers.py).  Covered: Dirichlet-Laplace and SSVS priors x {LDLT, SV}.
"""
from __future__ import annotations

import math

import numpy as np

# Philox stages
ST_COEF_Z, ST_IMPACT_Z, ST_SV_MIX_U, ST_SV_H_Z, ST_SV_SIGH, ST_SV_H0_Z, ST_LDLT_DIAG = 1, 2, 3, 4, 5, 6, 7
ST_PRIOR_COEF, ST_PRIOR_CONTEM = 16, 24
DBL_MIN = np.finfo(np.float64).tiny


class Draws:
    """Random variates at their Philox addresses, produced by the oracle's sampler library."""

    def __init__(self, lib, seed: int, unit: int):
        self.lib, self.seed, self.unit, self.iter = lib, int(seed) & 0xFFFFFFFF, int(unit), 0

    def _at(self, kind, stage, idx, p0=0.0, p1=0.0, p2=0.0):
        return self.lib.oracle_draw_at(kind, self.seed, self.unit, self.iter, stage, int(idx), float(p0), float(p1), float(p2))

    def normal_vec(self, stage, e0, n):
        return np.array([self._at(0, stage, e0 + i) for i in range(n)])

    def uniform_vec(self, stage, e0, n):
        return np.array([self._at(1, stage, e0 + i) for i in range(n)])

    def gamma(self, stage, idx, shape, scale):
        return self._at(2, stage, idx, shape, scale)

    def beta(self, stage, idx, a, b):
        return self._at(3, stage, idx, a, b)

    def invgauss(self, stage, idx, mean, shape):
        return self._at(4, stage, idx, mean, shape)

    def gig(self, stage, idx, lam, psi, chi):
        return self._at(5, stage, idx, lam, psi, chi)

    def uniform(self, stage, idx):
        return self._at(6, stage, idx)


def cut_param(x):  # draw.h:99-103
    return DBL_MIN if (x < DBL_MIN or math.isnan(x)) else x


def build_inv_lower(dim, lower_vec):  # draw.h:124-132
    res = np.eye(dim)
    i0 = 0
    for i in range(1, dim):
        res[i, :i] = lower_vec[i0:i0 + i]
        i0 += i
    return res


def cat_rand(weight, u):
    """discrete_distribution on normalised probabilities (common.h:577-580): inverse CDF with the injected uniform."""
    cum = np.cumsum(weight / weight.sum())
    idx = int(np.searchsorted(cum, u, side="right"))
    return min(idx, len(weight) - 1)


class MirrorChain:
    def __init__(self, kw: dict, chain: int, lib, window: int = 0):
        self.x = np.asarray(kw["x"], dtype=np.float64)
        self.y = np.asarray(kw["y"], dtype=np.float64)
        self.include_mean = bool(kw["include_mean"])
        self.n, self.dim = self.y.shape
        self.dim_design = self.x.shape[1]
        k = self.dim
        self.num_lowerchol = k * (k - 1) // 2
        self.num_coef = k * self.dim_design
        self.num_alpha = self.num_coef - k if self.include_mean else self.num_coef
        self.nrow = self.num_alpha // k
        self.prior = int(kw["prior_type"])
        self.ggl = bool(kw.get("is_group", True))
        self.grp_id = np.asarray(kw["grp_id"]).astype(int)
        self.grp_vec = np.asarray(kw["grp_mat"]).reshape(-1, order="F").astype(int)  # tri.h:42
        self.own_id = set(int(v) for v in np.asarray(kw["own_id"]).reshape(-1))
        cov, icpt, init, pp = kw["param_cov"], kw["param_intercept"], kw["param_init"][chain], kw["param_prior"]
        self.pp = pp
        self.sig_shp = np.asarray(cov["shape"], dtype=np.float64)
        self.sig_scl = np.asarray(cov["scale"], dtype=np.float64)
        self.sv = "initial_mean" in cov  # src/triangular.cpp:33
        seeds = np.asarray(kw["seed_chain"]).reshape(-1, int(kw["num_chains"]))
        unit = window * int(kw["num_chains"]) + chain
        self.rng = Draws(lib, int(seeds[window, chain]), unit)
        # McmcTriangular ctor, tri.h:40-71
        self.coef_mat = np.array(init["init_coef"], dtype=np.float64, order="F").reshape(self.dim_design, k, order="F")
        self.contem_coef = np.array(init["init_contem"], dtype=np.float64).reshape(-1)
        self.prior_alpha_mean = np.zeros(self.num_coef)
        self.prior_alpha_prec = np.zeros(self.num_coef)
        self.alpha_penalty = np.zeros(self.num_alpha)
        self.prior_chol_mean = np.zeros(self.num_lowerchol)
        self.prior_chol_prec = np.ones(self.num_lowerchol)
        self.sparse_coef = np.zeros((self.dim_design, k))
        self.sparse_contem = np.zeros(self.num_lowerchol)
        self.chol_lower = build_inv_lower(k, self.contem_coef)
        self.latent_innov = self.y - self.x @ self.coef_mat
        self.sqrt_sv = np.zeros((self.n, k))
        if self.include_mean:
            self.prior_alpha_mean[-k:] = np.asarray(icpt["mean_non"], dtype=np.float64)
            self.prior_alpha_prec[-k:] = 1.0 / (float(icpt["sd_non"]) * np.ones(k)) ** 2
        self._sync_coef_vec()
        if self.sv:
            self.lvol_draw = np.array(init["lvol"], dtype=np.float64).reshape(self.n, k, order="F")
            self.lvol_init = np.array(init["lvol_init"], dtype=np.float64)
            self.lvol_sig = np.array(init["lvol_sig"], dtype=np.float64)
            self.prior_init_mean = np.asarray(cov["initial_mean"], dtype=np.float64)
            self.prior_init_prec = np.asarray(cov["initial_prec"], dtype=np.float64).reshape(k, k)
        else:
            self.diag_vec = np.array(init["init_diag"], dtype=np.float64)
        G, N, q = len(self.grp_id), self.num_alpha, self.num_lowerchol
        if self.prior == 2:  # McmcSsvs ctor tri.h:563-577
            self.coef_grid, self.contem_grid = int(pp["coef_grid"]), int(pp["chol_grid"])
            self.coef_dummy = np.array(init["init_coef_dummy"], dtype=np.float64)
            self.coef_weight = np.array(init["coef_mixture"], dtype=np.float64)
            self.contem_dummy = np.ones(q)
            self.contem_weight = np.array(init["chol_mixture"], dtype=np.float64)
            self.coef_slab = np.array(init["coef_slab"], dtype=np.float64)
            self.spike_scl = float(init["coef_spike_scl"])
            self.contem_spike_scl = float(init["coef_spike_scl"])  # tri.h:569: initialised from the COEF spike scale
            self.contem_slab = np.array(init["contem_slab"], dtype=np.float64)
            self.slab_weight = np.ones(N)
        elif self.prior == 6:  # McmcDl ctor tri.h:852-863
            self.dir_concen = 0.0
            self.contem_dir_concen = 0.0
            self.local_lev = np.array(init["local_sparsity"], dtype=np.float64)
            self.group_lev = np.zeros(G)
            self.global_lev = float(init["global_sparsity"]) if self.ggl else 1.0
            self.latent_local = np.zeros(N)
            self.coef_var = np.zeros(N)
            self.contem_local_lev = np.array(init["contem_local_sparsity"], dtype=np.float64)
            self.contem_global_lev = np.array(init["contem_global_sparsity"], dtype=np.float64).reshape(-1)
            self.latent_contem_local = np.zeros(q)
        else:
            raise NotImplementedError("the NumPy mirror covers SSVS (2) and DL (6)")

    def _sync_coef_vec(self):
        self.coef_vec = np.zeros(self.num_coef)
        self.coef_vec[: self.num_alpha] = self.coef_mat[: self.nrow, :].reshape(-1, order="F")
        if self.include_mean:
            self.coef_vec[-self.dim:] = self.coef_mat[-1, :]

    # ------------------------------------------------------------------ draw.h:372-383
    def draw_coef(self, x, y, prior_mean, prior_prec, z):
        prec = np.diag(prior_prec) + x.T @ x
        L = np.linalg.cholesky(prec)
        post_mean = np.linalg.solve(prec, prior_prec * prior_mean + x.T @ y)
        return post_mean + np.linalg.solve(L.T, z)

    # ------------------------------------------------------------------ the sweep, tri.h:102-115
    def sweep(self, it: int):
        self.rng.iter = it
        self.updateCoefPrec()
        self.updatePenalty()
        self.updateSv()
        self.updateCoef()
        self.updateImpactPrec()
        self.updateLatent()
        self.updateImpact()
        self.updateChol()
        self.updateState()

    def updatePenalty(self):  # tri.h:254-262
        for i in range(self.num_alpha):
            self.alpha_penalty[i] = 0.0 if self.grp_vec[i] in self.own_id else 1.0

    def updateSv(self):  # tri.h:372, 409
        if self.sv:
            self.sqrt_sv = np.exp(self.lvol_draw / 2)
        else:
            self.sqrt_sv = np.tile(np.sqrt(self.diag_vec)[None, :], (self.n, 1))

    def updateCoef(self):  # tri.h:281-316
        k, d, nrow = self.dim, self.dim_design, self.nrow
        xnorm = (self.x ** 2).sum(axis=0)
        for j in range(k):
            self.coef_mat[:, j] = 0.0
            chol_lower_j = self.chol_lower[j:, :]
            sqrt_sv_j = self.sqrt_sv[:, j:]
            design = np.kron(chol_lower_j[:, j:j + 1], self.x) / sqrt_sv_j.reshape(-1, order="F")[:, None]
            resp = (((self.y - self.x @ self.coef_mat) @ chol_lower_j.T) / sqrt_sv_j).reshape(-1, order="F")
            penalty_j = np.zeros(d)
            if self.include_mean:
                mean_j = np.concatenate([self.prior_alpha_mean[j * nrow:(j + 1) * nrow], [self.prior_alpha_mean[-k:][j]]])
                prec_j = np.concatenate([self.prior_alpha_prec[j * nrow:(j + 1) * nrow], [self.prior_alpha_prec[-k:][j]]])
                penalty_j[:nrow] = self.alpha_penalty[j * nrow:(j + 1) * nrow]
            else:
                mean_j = self.prior_alpha_mean[d * j:d * (j + 1)]
                prec_j = self.prior_alpha_prec[d * j:d * (j + 1)]
                penalty_j = self.alpha_penalty[d * j:d * (j + 1)]
            z = self.rng.normal_vec(ST_COEF_Z, j * d, d)
            self.coef_mat[:, j] = self.draw_coef(design, resp, mean_j, prec_j, z)
            self._sync_coef_vec()
            # draw_mn_savs draw.h:417-430
            coef = self.coef_mat[:, j]
            with np.errstate(divide="ignore", invalid="ignore"):
                pen = penalty_j / coef ** 2
                abs_fit = np.abs(coef) * xnorm
                self.sparse_coef[:, j] = coef * np.maximum(abs_fit - pen, 0.0) / abs_fit

    def updateLatent(self):  # tri.h:342
        self.latent_innov = self.y - self.x @ self.coef_mat

    def updateImpact(self):  # tri.h:322-336
        for j in range(1, self.dim):
            resp = self.latent_innov[:, j] / self.sqrt_sv[:, j]
            design = self.latent_innov[:, :j] / self.sqrt_sv[:, j][:, None]
            cid = j * (j - 1) // 2
            z = self.rng.normal_vec(ST_IMPACT_Z, cid, j)
            a = self.draw_coef(design, resp, self.prior_chol_mean[cid:cid + j], self.prior_chol_prec[cid:cid + j], z)
            self.contem_coef[cid:cid + j] = a
            # draw_savs draw.h:398-415
            with np.errstate(divide="ignore", invalid="ignore"):
                pen = 1.0 / a ** 2
                x_norm = (self.latent_innov[:, :j] ** 2).sum(axis=0)
                abs_fit = np.abs(a) * x_norm
                self.sparse_contem[cid:cid + j] = a * np.maximum(abs_fit - pen, 0.0) / abs_fit

    def updateChol(self):  # tri.h:348
        self.chol_lower = build_inv_lower(self.dim, self.contem_coef)

    def updateState(self):
        n, k = self.n, self.dim
        if not self.sv:  # reg_ldlt_diag draw.h:1244-1259 (num_design / 2: integer division)
            ortho = self.latent_innov @ self.chol_lower.T
            for i in range(k):
                g = self.rng.gamma(ST_LDLT_DIAG, i, self.sig_shp[i] + n // 2, 1.0 / (self.sig_scl[i] + (ortho[:, i] ** 2).sum() / 2))
                self.diag_vec[i] = 1.0 / g
            return
        ortho = np.log((self.latent_innov @ self.chol_lower.T) ** 2 + .0001)  # tri.h:401-402
        for i in range(k):
            self.lvol_draw[:, i] = self.varsv_ht(i, self.lvol_draw[:, i].copy(), self.lvol_init[i], self.lvol_sig[i], ortho[:, i])
        # varsv_sigh draw.h:498-512: .sum() runs over the WHOLE matrix; num_design / 2 integer division
        h_slide = np.vstack([self.lvol_init[None, :], self.lvol_draw[:-1, :]])
        ss = ((self.lvol_draw - h_slide) ** 2).sum()
        for i in range(k):
            self.lvol_sig[i] = 1.0 / self.rng.gamma(ST_SV_SIGH, i, self.sig_shp[i] + n // 2, 1.0 / (self.sig_scl[i] + ss / 2))
        # varsv_h0 draw.h:523-537
        z = self.rng.normal_vec(ST_SV_H0_Z, 0, k)
        h_diagprec = np.diag(1.0 / self.lvol_sig)
        post = self.prior_init_prec + h_diagprec
        L = np.linalg.cholesky(post)
        mean = np.linalg.solve(post, self.prior_init_prec @ self.prior_init_mean + h_diagprec @ self.lvol_draw[0, :])
        self.lvol_init = mean + np.linalg.solve(L.T, z)

    def varsv_ht(self, series, sv_vec, init_sv, sv_sig, latent_vec):  # draw.h:440-488, dense
        n = self.n
        pj = np.array([0.0073, 0.10556, 0.00002, 0.04395, 0.34001, 0.24566, 0.2575])
        muj = np.array([-10.12999, -3.97281, -8.56686, 2.77786, 0.61942, 1.79518, -1.08819]) - 1.2704
        sigj = np.array([5.79596, 2.61369, 5.17950, 0.16735, 0.64009, 0.34023, 1.26261])
        sdj = np.sqrt(sigj)
        base = series * 2 * ((n + 1) // 2)  # address of (series, t): see DESIGN.md section 5
        inv_method = self.rng.uniform_vec(ST_SV_MIX_U, base, n)
        pdf = np.exp(-((latent_vec[:, None] - sv_vec[:, None] - muj[None, :]) / sdj[None, :]) ** 2 / 2) * pj[None, :] / (sdj[None, :] * math.sqrt(2 * math.pi))
        pdf = pdf / pdf.sum(axis=1, keepdims=True)
        cumsum = np.cumsum(pdf, axis=1)
        binom_latent = 7 - (inv_method[:, None] < cumsum).sum(axis=1)
        binom_latent = np.minimum(binom_latent, 6)  # (rounding guard: cumsum[-1] may be 1 - ulp)
        ds = muj[binom_latent]
        inv_sig_s = np.diag(1.0 / sigj[binom_latent])
        diff_mat = np.eye(n)
        for i in range(n - 1):
            diff_mat[i + 1, i] = -1.0
        HtH = diff_mat.T @ diff_mat
        res = self.rng.normal_vec(ST_SV_H_Z, base, n)
        post_sig = HtH / sv_sig + inv_sig_s
        L = np.linalg.cholesky(post_sig)
        post_mean = np.linalg.solve(post_sig, HtH @ (init_sv * np.ones(n)) / sv_sig + inv_sig_s @ (latent_vec - ds))
        return post_mean + np.linalg.solve(L.T, res)

    # ------------------------------------------------------------------ priors
    def updateCoefPrec(self):
        N, P0 = self.num_alpha, ST_PRIOR_COEF
        coef = self.coef_vec[:N]
        if self.prior == 2:  # McmcSsvs::updateCoefPrec tri.h:599-616
            pp = self.pp
            for i in range(N):  # ssvs_local_slab draw.h:336-345
                g = self.rng.gamma(P0 + 0, i, float(pp["coef_slab_shape"]) + .5,
                                   1.0 / (float(pp["coef_slab_scl"]) + coef[i] * coef[i] / (self.coef_dummy[i] + (1 - self.coef_dummy[i]) * self.spike_scl)))
                self.coef_slab[i] = math.sqrt(1.0 / g)
            for j, gid in enumerate(self.grp_id):
                self.slab_weight = np.where(self.grp_vec == gid, self.coef_weight[j], self.slab_weight)
            self.spike_scl = self.ssvs_scl_griddy(P0 + 1, self.coef_grid, coef, self.coef_slab)
            self.coef_dummy = self.ssvs_dummy(P0 + 2, coef, self.coef_slab, self.spike_scl * self.coef_slab, self.slab_weight)
            s1, s2 = np.asarray(pp["coef_s1"], dtype=np.float64), np.asarray(pp["coef_s2"], dtype=np.float64)
            for j, gid in enumerate(self.grp_id):  # ssvs_mn_weight draw.h:308-325
                sel = self.grp_vec == gid
                mn = self.coef_dummy[sel]
                self.coef_weight[j] = self.rng.beta(P0 + 3, j, s1[j] + mn.sum(), s2[j] + mn.size - mn.sum())
            # tri.h:615 (the sd, not its square)
            self.prior_alpha_prec[:N] = 1.0 / (self.spike_scl * (1 - self.coef_dummy) * self.coef_slab + self.coef_dummy * self.coef_slab)
        else:  # McmcDl::updateCoefPrec tri.h:885-901
            shape, scl, grid = float(self.pp["shape"]), float(self.pp["scale"]), int(self.pp["grid_size"])
            for j, gid in enumerate(self.grp_id):  # dl_mn_sparsity (first overload) draw.h:808-834
                sel = self.grp_vec == gid
                mn_scl = np.abs(coef[sel]) / (self.global_lev * self.local_lev[sel])
                self.group_lev[j] = cut_param(1.0 / self.rng.gamma(P0 + 0, j, shape + int(sel.sum()), 1.0 / (scl + mn_scl.sum())))
            for j, gid in enumerate(self.grp_id):
                self.coef_var = np.where(self.grp_vec == gid, self.group_lev[j], self.coef_var)
            self.dir_concen = self.dl_dir_griddy(P0 + 1, grid, self.local_lev, self.global_lev)
            self.local_lev = self.dl_local_sparsity(P0 + 2, self.dir_concen, coef / self.coef_var)
            if self.ggl:
                self.global_lev = self.dl_global_sparsity(P0 + 3, self.local_lev * self.coef_var, self.dir_concen, coef)
            self.latent_local = self.dl_latent(P0 + 4, self.global_lev * self.local_lev * self.coef_var, coef)
            self.prior_alpha_prec[:N] = 1.0 / ((self.global_lev * self.local_lev * self.coef_var) ** 2 * self.latent_local)

    def updateImpactPrec(self):
        q, P1 = self.num_lowerchol, ST_PRIOR_CONTEM
        a = self.contem_coef
        if self.prior == 2:  # tri.h:617-623
            pp = self.pp
            for i in range(q):
                g = self.rng.gamma(P1 + 0, i, float(pp["chol_slab_shape"]) + .5,
                                   1.0 / (float(pp["chol_slab_scl"]) + a[i] * a[i] / (self.contem_dummy[i] + (1 - self.contem_dummy[i]) * self.contem_spike_scl)))
                self.contem_slab[i] = math.sqrt(1.0 / g)
            self.contem_spike_scl = self.ssvs_scl_griddy(P1 + 1, self.contem_grid, a, self.contem_slab)
            self.contem_dummy = self.ssvs_dummy(P1 + 2, a, self.contem_slab, self.contem_spike_scl * self.contem_slab, self.contem_weight)
            s1 = float(pp["chol_s1"]) + self.contem_dummy.sum()  # ssvs_weight draw.h:290-297
            s2 = float(pp["chol_s2"]) + q - self.contem_dummy.sum()
            for i in range(q):
                self.contem_weight[i] = self.rng.beta(P1 + 3, i, s1, s2)
            sd = (1 - self.contem_dummy) * (self.contem_spike_scl * self.contem_slab) + self.contem_dummy * self.contem_slab  # draw.h:112-116
            self.prior_chol_prec = 1.0 / sd ** 2
        else:  # tri.h:902-908
            grid = int(self.pp["grid_size"])
            self.contem_dir_concen = self.dl_dir_griddy(P1 + 0, grid, self.contem_local_lev, self.contem_global_lev[0])
            self.contem_local_lev = self.dl_local_sparsity(P1 + 1, self.contem_dir_concen, a)
            self.contem_global_lev[0] = self.dl_global_sparsity(P1 + 2, self.contem_local_lev, self.contem_dir_concen, a)
            self.latent_contem_local = self.dl_latent(P1 + 3, self.contem_local_lev, a)
            self.prior_chol_prec = 1.0 / ((self.contem_global_lev[0] * self.contem_local_lev) ** 2 * self.latent_contem_local)

    def ssvs_scl_griddy(self, stage, grid_size, coef, slab):  # draw.h:347-361
        grid = np.linspace(0.0, 1.0, grid_size + 2)[1:grid_size + 1]
        log_wt = np.array([-((coef / slab) ** 2).sum() / (2 * c * c) - coef.size * math.log(c) for c in grid])
        weight = np.exp(log_wt - log_wt.max())
        weight /= weight.sum()
        return float(grid[cat_rand(weight, self.rng.uniform(stage, 0))])

    def ssvs_dummy(self, stage, param, sd_numer, sd_denom, slab_weight):  # draw.h:269-281
        u1 = -param ** 2 / (2 * sd_numer ** 2)
        u2 = -param ** 2 / (2 * sd_denom ** 2)
        mx = np.maximum(u1, u2)
        e1 = slab_weight * np.exp(u1 - mx) / sd_numer
        e2 = (1 - slab_weight) * np.exp(u2 - mx) / sd_denom
        out = np.empty(param.size)
        for i in range(param.size):
            out[i] = 1.0 if self.rng.uniform(stage, i) < e1[i] / (e1[i] + e2[i]) else 0.0
        return out

    def dl_dir_griddy(self, stage, grid_size, local, glob):  # draw.h:866-890
        size = local.size
        lo0 = float(1 // size)  # `1 / local_param.size()`: INTEGER division (0 unless size == 1)
        grid = np.linspace(lo0, .5, grid_size) if lo0 < .5 else np.linspace(.5, lo0, grid_size)
        log_wt = np.empty(grid_size)
        with np.errstate(divide="ignore", invalid="ignore"):
            for i, c in enumerate(grid):
                log_wt[i] = c * (size * (math.log(glob) - math.log(2.0)) + np.log(local).sum()) - size * (math.lgamma(c) if c > 0 else math.inf)
        weight = np.exp(log_wt - np.nanmax(log_wt))
        weight = np.where(np.isnan(weight), 0.0, weight)
        weight /= weight.sum()
        return float(grid[cat_rand(weight, self.rng.uniform(stage, 0))])

    def dl_local_sparsity(self, stage, dir_concen, coef):  # draw.h:778-785
        out = np.array([cut_param(self.rng.gig(stage, i, dir_concen - 1, 1.0, 2 * abs(coef[i]))) for i in range(coef.size)])
        return out / out.sum()

    def dl_global_sparsity(self, stage, local, dir_concen, coef):  # draw.h:793-799
        return cut_param(self.rng.gig(stage, 0, coef.size * (dir_concen - 1), 1.0, 2 * (np.abs(coef) / local).sum()))

    def dl_latent(self, stage, local, coef):  # draw.h:758-770
        return np.array([cut_param(1.0 / self.rng.invgauss(stage, i, local[i] / abs(coef[i]), 1.0)) for i in range(coef.size)])

    # ------------------------------------------------------------------ views in the layout of the engines' get_state
    def state(self, name):
        if name == "coef_mat":
            return self.coef_mat.reshape(-1, order="F")
        if name == "sparse_coef":
            return self.sparse_coef.reshape(-1, order="F")
        if name == "chol_lower":
            return self.chol_lower.reshape(-1, order="F")
        if name == "latent_innov":
            return self.latent_innov.reshape(-1, order="F")
        if name == "lvol_draw":
            return self.lvol_draw.reshape(-1, order="F")
        if name == "sqrt_sv":
            return self.sqrt_sv.reshape(-1, order="F")
        if name in ("spike_scl", "contem_spike_scl", "global_lev", "dir_concen", "contem_dir_concen"):
            return np.array([getattr(self, name)])
        if name == "contem_global_lev":
            return np.array([self.contem_global_lev[0]])
        return np.asarray(getattr(self, name), dtype=np.float64).reshape(-1)

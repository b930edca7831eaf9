"""Bayesian VAR / VHAR front end on the CUDA Gibbs engine.

``VarBayes`` and ``VharBayes`` keep the constructor arguments, attributes and method names of the
reference's Python classes (python/src/bvhar/model/_bayes.py:17-181 ``_AutoregBayes``, :183-664
``VarBayes``, :666-1113 ``VharBayes``; R: ``var_bayes()`` R/var-bayes.R:81, ``vhar_bayes()``
R/vhar-bayes.R, ``forecast_roll()/forecast_expand()`` R/summary-forecast.R:412-588).  What differs is
underneath: every chain of ``fit()`` and every (window, chain) refit of ``roll_forecast()`` /
``expand_forecast()`` runs as one batch of units on the GPU.

Draws are kept as ``{record name: ndarray[chain, draw, column]}`` in ``param_`` (the reference builds
a pandas frame of the same columns, _misc.py:122-148); ``to_frame()`` produces that frame on demand.
"""
from __future__ import annotations

import warnings
from math import floor
from typing import Dict, List, Optional

import numpy as np

from . import _abi
from ._abi import REC_ALL, REC_COEF, REC_LVOL, REC_LVOL_LAST, REC_SPARSE, PackedFit
from ._design import build_design, build_grpmat, build_response, build_vhar, build_vhar_design
from ._engine import CudaFit, forecast_density, outforecast
from ._spec import (PRIOR_TYPE, DlConfig, GdpConfig, HorseshoeConfig, InterceptConfig, LdltConfig, MinnesotaConfig,
                    NgConfig, SsvsConfig, SvConfig, _BayesConfig, make_inits)


def _check_np(data) -> np.ndarray:
    arr = np.asarray(data.values if hasattr(data, "values") else data)
    if arr.ndim != 2:
        raise ValueError("Numpy array must be 2-dim.")
    if not np.issubdtype(arr.dtype, np.number) or np.issubdtype(arr.dtype, np.bool_):
        raise ValueError("All elements should be numeric.")
    return np.asfortranarray(arr, dtype=np.float64)


class _AutoregBayes:
    def __init__(self, data, lag, p, n_chain=1, n_iter=1000, n_burn=None, n_thin=1, bayes_config=None, cov_config=None,
                 intercept_config=None, fit_intercept=True, minnesota="longrun", ggl=True, verbose=False, seed=None,
                 device: int = 0, record_mask: Optional[int] = None):
        self.y_ = _check_np(data)
        self.n_features_in_ = self.y_.shape[1]
        self.p_ = p
        self.lag_ = lag
        if self.y_.shape[0] <= self.lag_:
            raise ValueError(f"'data' rows must be larger than 'lag' = {self.lag_}")
        self.chains_ = int(n_chain)
        self.iter_ = int(n_iter)
        self.burn_ = int(floor(n_iter / 2) if n_burn is None else n_burn)
        self.thin_ = int(n_thin)
        self.fit_intercept = fit_intercept
        self.group_ = build_grpmat(self.p_, self.n_features_in_, minnesota)
        self._group_id = np.array(list(dict.fromkeys(self.group_.flatten(order="F").tolist())), dtype=np.int32)
        self._own_id = None
        self._cross_id = None
        self._ggl = ggl
        self._verbose = verbose
        self._device = device
        self._record_mask = record_mask
        self._rng = np.random.default_rng(seed)
        self.cov_spec_ = LdltConfig() if cov_config is None else cov_config
        self.spec_ = SsvsConfig() if bayes_config is None else bayes_config
        self.intercept_spec_ = InterceptConfig() if intercept_config is None else intercept_config
        self._prior_type = PRIOR_TYPE.get(self.spec_.prior)
        self.is_fitted_ = False
        self.coef_ = None
        self.intercept_ = None
        self.param_names_ = None
        self.param_: Optional[Dict[str, np.ndarray]] = None
        self.sparse_coef_ = None
        self.summary_ = None
        self.init_ = None
        self._fit_obj = None

    # -- shared pieces ----------------------------------------------------------------------
    def _validate(self):
        if not isinstance(self.cov_spec_, LdltConfig):
            raise TypeError("`cov_config` should be `LdltConfig` or `SvConfig`.")
        if not isinstance(self.intercept_spec_, InterceptConfig):
            raise TypeError("`intercept_config` should be `InterceptConfig` when 'fit_intercept' is True.")
        if not isinstance(self.spec_, _BayesConfig):
            raise TypeError("`bayes_spec` should be the derived class of `_BayesConfig`.")
        self.cov_spec_.update(self.n_features_in_)
        self.intercept_spec_.update(self.n_features_in_)
        if isinstance(self.spec_, SsvsConfig):
            self.spec_.update(self._group_id, self._own_id, self._cross_id)
        elif isinstance(self.spec_, MinnesotaConfig):
            self.spec_.update(self.y_, self.p_, self.n_features_in_)

    @property
    def _is_sv(self) -> bool:
        return isinstance(self.cov_spec_, SvConfig)

    def _make_inits(self):
        n_design = self.design_.shape[1]
        self.init_ = make_inits(self.spec_.prior, self._is_sv, self.n_features_in_, n_design, self.response_.shape[0],
                                len(self._group_id), self.chains_, self._rng, include_mean=self.fit_intercept)

    def _seeds(self, *shape):
        return self._rng.integers(1, np.iinfo(np.int32).max, size=shape)

    def _default_record_mask(self) -> int:
        """Record policy when the caller gives none: everything the reference's ``returnRecords()`` returns - except that
        batches of >= 64 stochastic-volatility chains keep the log-volatilities of the LAST time point (``h_last_record``,
        all the forecasters read: config.h:998-1001) instead of the full ``h_record`` (n x k doubles per draw per chain:
        0.94 TB at BASELINE's C3).  ``record_mask=REC_ALL`` brings the full ``h_record`` back."""
        if self._record_mask is not None:
            return int(self._record_mask)
        if self._is_sv and self.chains_ >= 64:
            return (REC_ALL & ~REC_LVOL) | REC_LVOL_LAST
        return REC_ALL

    def _pack(self, x, y, seeds, win_row0=None, win_nrow=None, record_mask=None, lvol_from_init=False, unit_id_base=0) -> PackedFit:
        return PackedFit(self.chains_, self.iter_, self.burn_, self.thin_, x, y, self.cov_spec_.to_dict(),
                         self.spec_.to_dict(), self.intercept_spec_.to_dict(), self.init_, int(self._prior_type),
                         self._group_id, self._own_id, self._cross_id, self.group_, self.fit_intercept, seeds,
                         is_group=self._ggl, win_row0=win_row0, win_nrow=win_nrow,
                         record_mask=self._default_record_mask() if record_mask is None else record_mask, device=self._device,
                         lvol_from_init=lvol_from_init, unit_id_base=unit_id_base)

    def _coef_name(self) -> str:
        return "alpha"

    def _fit_sharded(self, devices):
        """Chains split over several engines (one per entry of ``devices``, normally one per GPU; the same ordinal may
        appear twice), each driven by its own host thread.  No collective: the Philox keys are global
        (``unit_id_base``), so the draws equal the single-engine run bit for bit; the records are concatenated."""
        import threading
        from ._partition import shard_fit_kwargs, shard_units
        seeds = self._seeds(1, self.chains_)
        kw = dict(num_chains=self.chains_, num_iter=self.iter_, num_burn=self.burn_, thin=self.thin_, x=self.design_,
                  y=self.response_, param_cov=self.cov_spec_.to_dict(), param_prior=self.spec_.to_dict(),
                  param_intercept=self.intercept_spec_.to_dict(), param_init=self.init_, prior_type=int(self._prior_type),
                  grp_id=self._group_id, own_id=self._own_id, cross_id=self._cross_id, grp_mat=self.group_,
                  include_mean=self.fit_intercept, seed_chain=seeds, is_group=self._ggl, record_mask=self._default_record_mask())
        world = len(devices)
        fits, errs = [None] * world, [None] * world

        def work(r):
            try:
                skw = shard_fit_kwargs(kw, shard_units(1, self.chains_, world, r))
                if skw is None:
                    return
                skw["device"] = int(devices[r])
                f = CudaFit(PackedFit(**skw))
                try:
                    f.run(None)
                except _abi.BvharCudaError as e:
                    if e.status != 3:
                        raise
                    warnings.warn(str(e))
                fits[r] = f
            except Exception as e:  # re-raised on the caller's thread
                errs[r] = e

        threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        for e in errs:
            if e is not None:
                raise e
        live = [f for f in fits if f is not None]
        names = live[0].record_names()
        recs = {nm: np.concatenate([f.record_all(nm, pinned=False) for f in live], axis=0) for nm in names}
        for f in live:
            f.close()
        return names, recs

    def fit(self, devices=None, summaries: bool = False, level: float = .05):
        """Conduct MCMC on the GPU and compute posterior means.  ``devices``: optional list of CUDA ordinals to split
        the chains over (one engine and one host thread per entry).  ``summaries=True`` also leaves, in ``summary_``, the
        posterior summaries the reference's users compute on the returned draws (posterior::summarise_draws: mean, sd,
        ``level`` credible interval, split-R-hat, ESS, plus the PIP), reduced on the device."""
        if devices is not None and len(devices) > 1:
            names, recs = self._fit_sharded(list(devices))
            k = self.n_features_in_
            # NaN-skipping means, as the device reduction of the single-engine path (record_mean): a chain with a
            # numerical flag carries NaNs and must not poison the posterior mean on one path only
            self.coef_ = np.nanmean(recs["alpha_record"], axis=(0, 1)).reshape(k, -1).T
            if "alpha_sparse_record" in recs:
                self.sparse_coef_ = np.nanmean(recs["alpha_sparse_record"], axis=(0, 1)).reshape(k, -1).T
            if self.fit_intercept:
                self.intercept_ = np.nanmean(recs["c_record"], axis=(0, 1)).reshape(k)
                self.coef_ = np.concatenate([self.coef_, self.intercept_[None, :]], axis=0)
                if self.sparse_coef_ is not None:
                    self.sparse_coef_ = np.concatenate([self.sparse_coef_, np.nanmean(recs["c_sparse_record"], axis=(0, 1))[None, :]], axis=0)
            self.param_ = recs
            self.param_names_ = [nm.replace("_record", "") for nm in names]
            self.is_fitted_ = True
            return self
        packed = self._pack(self.design_, self.response_, self._seeds(1, self.chains_))
        fit = CudaFit(packed)
        cb = None
        if self._verbose:
            cb = lambda done, total: print(f"[{self.chains_} chains] {done} / {total}", flush=True) or False
        try:
            fit.run(cb)
        except _abi.BvharCudaError as e:
            if e.status != 3:
                raise
            warnings.warn(str(e))
        names = fit.record_names()
        recs = {nm: fit.record_all(nm) for nm in names}  # [chain, draw, column]
        k = self.n_features_in_
        # posterior means reduced on the device (no host pass over the draws)
        self.coef_ = fit.record_mean("alpha_record").reshape(k, -1).T
        if "alpha_sparse_record" in recs:
            self.sparse_coef_ = fit.record_mean("alpha_sparse_record").reshape(k, -1).T
        if self.fit_intercept:
            self.intercept_ = fit.record_mean("c_record").reshape(k)
            self.coef_ = np.concatenate([self.coef_, self.intercept_[None, :]], axis=0)
            if self.sparse_coef_ is not None:
                self.sparse_coef_ = np.concatenate([self.sparse_coef_, fit.record_mean("c_sparse_record")[None, :]], axis=0)
        if summaries:  # mean / sd / credible interval / split-R-hat / ESS / pip of every column, reduced on the device
            self.summary_ = {nm.replace("_record", ""): fit.record_summary(nm, level) for nm in names if recs[nm].shape[2] > 0 and recs[nm].shape[1] > 3}
        fit.close()
        self.param_ = recs
        self.param_names_ = [nm.replace("_record", "") for nm in names]
        self.is_fitted_ = True
        return self

    def to_frame(self):
        """Draws as the reference's wide pandas frame (one column per parameter, _misc.py:122-148)."""
        import pandas as pd
        frames = []
        tot = 0
        for c in range(self.chains_):
            cols = {}
            for nm, arr in self.param_.items():
                base = nm.replace("_record", "")
                if self._coef_name() == "phi":
                    base = base.replace("alpha", "phi")
                for j in range(arr.shape[2]):
                    cols[f"{base}[{j + 1}]"] = arr[c, :, j]
            df = pd.DataFrame(cols)
            df["_chain"] = c
            df["_iteration"] = np.arange(1, len(df) + 1)
            df["_draw"] = np.arange(tot + 1, tot + len(df) + 1)
            tot += len(df)
            frames.append(df)
        return pd.concat(frames, axis=0)

    def _chain_records(self, sparse=False) -> List[Dict[str, np.ndarray]]:
        if not self.is_fitted_:
            raise RuntimeError("call fit() first")
        return [{nm: arr[c] for nm, arr in self.param_.items()} for c in range(self.chains_)]

    @staticmethod
    def _summarise(dens: np.ndarray, dim: int, level: float, med: bool):
        step = dens.shape[0]
        draws = dens.reshape(step, -1, dim)  # [step, draw, variable]
        return {
            "forecast": np.median(draws, axis=1) if med else np.mean(draws, axis=1),
            "se": np.std(draws, axis=1, ddof=1),
            "lower": np.quantile(draws, level / 2, axis=1),
            "upper": np.quantile(draws, 1 - level / 2, axis=1),
        }

    def _har(self):
        return None

    @staticmethod
    def _ci_level(sparse):
        """``sparse`` given as a number = credible-interval level of the Select forecasters (R/forecast.R:423-428:
        ``ci_lev <- sparse; sparse <- FALSE``); returns (sparse flag, level)."""
        if isinstance(sparse, (bool, np.bool_)):
            return bool(sparse), 0.0
        if isinstance(sparse, (int, float, np.integer, np.floating)):
            return False, float(sparse)
        raise ValueError("'sparse' must be a bool or a credible-interval level")

    def predict(self, n_ahead: int, level=.05, stable=False, sparse=False, med=False, sv=True):
        """'n_ahead'-step density forecast from the stored draws (forecaster.h:494-568).  ``sparse``: ``True`` uses the
        SAVS draws; a number uses the full draws restricted by the activity graph of that credible level
        (``McmcVarSelectForecaster`` / ``McmcVharSelectForecaster``, forecaster.h:348-416)."""
        sparse, ci_lev = self._ci_level(sparse)
        recs = self._chain_records()
        seeds = self._seeds(self.chains_)
        if stable:  # the stability filter (bvhar_cuda_is_stable) may keep a different number of draws per chain: one call per chain
            dens = [forecast_density([rec], [self.y_], n_ahead, self.lag_, self.fit_intercept, self._is_sv, [seeds[c]],
                                     har_trans=self._har(), sv=sv, stable=True, sparse=sparse, device=self._device,
                                     unit_id_base=c, level=ci_lev)[0][0] for c, rec in enumerate(recs)]
        else:
            dens, _ = forecast_density(recs, [self.y_] * self.chains_, n_ahead, self.lag_, self.fit_intercept, self._is_sv, seeds,
                                       har_trans=self._har(), sv=sv, sparse=sparse, device=self._device, level=ci_lev)
        k = self.n_features_in_
        allchains = np.concatenate([d.reshape(n_ahead, -1, k) for d in dens], axis=1).reshape(n_ahead, -1)
        return self._summarise(allchains, k, level, med)

    def spillover(self, n_ahead: int = 10, level=.05, sparse=True):
        """Spillover densities from the stored draws of all chains (the reference's ``spillover()``, _bayes.py:611-641 ->
        ``LdltSpillover`` / ``SvSpillover`` = ``McmcSpilloverRun``, spillover.h:246-260), computed on the device."""
        if not self.is_fitted_:
            raise RuntimeError("call fit() first")
        from ._engine import spillover as _spillover
        k = self.n_features_in_
        sfx = "_sparse_record" if sparse else "_record"
        if sparse and "alpha_sparse_record" not in self.param_:
            raise ValueError("sparse draws were not recorded")
        coef = np.concatenate(list(self.param_["alpha" + sfx]), axis=0)   # concat_params(..., chain=False): every chain's draws
        contem = np.concatenate(list(self.param_["a" + sfx]), axis=0) if k > 1 else np.zeros((coef.shape[0], 1))
        nrow = coef.shape[1] // k
        if self._is_sv:
            if "h_record" not in self.param_:
                raise ValueError("spillover() of an SV fit needs the full 'h_record'")
            sv = np.concatenate(list(self.param_["h_record"]), axis=0)
            kind, n_time = 2, sv.shape[1] // k
        else:
            sv = np.concatenate(list(self.param_["d_record"]), axis=0)
            kind, n_time = 0, 0
        out = _spillover(coef, contem, sv, k, nrow, self.lag_, n_ahead, har_trans=self._har(), sv_kind=kind, n_time=n_time, device=self._device)

        def tab(m, lv=None):  # process_dens_spillover, _misc.py:172-181
            st = np.stack(np.array_split(m, m.shape[1] // k, axis=1), axis=0)
            return {"mean": st.mean(axis=0), "se": st.std(axis=0), "lower": np.quantile(st, level / 2, axis=0),
                    "upper": np.quantile(st, 1 - level / 2, axis=0)}

        def vec(v):  # process_dens_vector_spillover, _misc.py:183-190
            r = v.reshape(-1, k)
            return {"mean": r.mean(axis=0), "se": r.std(axis=0), "lower": np.quantile(r, level / 2, axis=0),
                    "upper": np.quantile(r, 1 - level / 2, axis=0)}

        return {"connect": tab(out["connect"]), "net_pairwise": tab(out["net_pairwise"]), "tot": vec(out["tot"]),
                "to": vec(out["to"]), "from": vec(out["from"]), "net": vec(out["net"])}

    def dynamic_spillover(self, window: int, n_ahead: int = 10, level=.05, sparse=True):
        """Dynamic spillover (the reference's ``dynamic_spillover()``, _bayes.py:643-666).  LDLT models: rolling windows of
        ``window`` observations, every window refitted with ``chains_`` chains (``DynamicLdltSpillover``, spillover.h:272-443).
        SV models: one spillover per time point from the stored draws and their log-volatilities (``DynamicSvSpillover``,
        spillover.h:445-507).  Returns, per horizon, mean / se / lower / upper of tot, to, from and net."""
        if not self.is_fitted_:
            raise RuntimeError("call fit() first")
        from ._engine import dynamic_spillover as _dyn, spillover as _spillover
        k = self.n_features_in_
        if self._is_sv:
            sfx = "_sparse_record" if sparse else "_record"
            if "h_record" not in self.param_:
                raise ValueError("dynamic_spillover() of an SV fit needs the full 'h_record'")
            coef = np.concatenate(list(self.param_["alpha" + sfx]), axis=0)
            contem = np.concatenate(list(self.param_["a" + sfx]), axis=0) if k > 1 else np.zeros((coef.shape[0], 1))
            h = np.concatenate(list(self.param_["h_record"]), axis=0)
            n_time, S = h.shape[1] // k, coef.shape[0]
            out = _spillover(coef, contem, h, k, coef.shape[1] // k, self.lag_, n_ahead, har_trans=self._har(), sv_kind=3, n_time=n_time,
                             device=self._device)
            to = out["to"].reshape(n_time, S, k); frm = out["from"].reshape(n_time, S, k); tot = out["tot"].reshape(n_time, S, 1)
        else:
            H = self.y_.shape[0] - window + 1
            if H <= 0:
                raise ValueError("Window size is too large")  # spillover.h:316-318
            x_tot, y_tot = self._design_of(self.y_)
            n_w = window - self.lag_
            packed = self._pack(x_tot, y_tot, self._seeds(H, self.chains_), win_row0=list(range(H)), win_nrow=[n_w] * H, record_mask=0)
            to, frm, tot = _dyn(packed, n_ahead, self.lag_, har_trans=self._har(), sparse=sparse)
            to = to.reshape(H, -1, k); frm = frm.reshape(H, -1, k); tot = tot.reshape(H, -1, 1)

        def dens(a):  # process_dens_list_spillover, _misc.py:195-209
            r = {"mean": a.mean(axis=1), "se": a.std(axis=1), "lower": np.quantile(a, level / 2, axis=1), "upper": np.quantile(a, 1 - level / 2, axis=1)}
            return {nm: (v[:, 0] if v.shape[1] == 1 else v) for nm, v in r.items()}

        return {"tot": dens(tot), "to": dens(to), "from": dens(frm), "net": dens(to - frm)}

    def _outforecast(self, n_ahead: int, test, expanding: bool, level, stable, sparse, med, sv):
        """forecast_roll / forecast_expand: the draws of ``_outforecast_draws`` summarised per window
        (process_forecast_draws, R/misc-r.R:186-268; _bayes.py:500-512)."""
        k = self.n_features_in_
        per_window, lpls = self._outforecast_draws(n_ahead, test, expanding, stable, sparse, sv)
        out = {"forecast": [], "se": [], "lower": [], "upper": []}
        for pw in per_window:
            s = self._summarise(pw.reshape(1, -1), k, level, med)  # all chains' draws of the window pooled
            for key in out:
                out[key].append(s[key])
        res = {key: np.concatenate(v, axis=0) for key, v in out.items()}
        if lpls is not None:
            l = np.stack(lpls)  # [window, chain]
            res["lpl"] = np.median(l, axis=1) if med else np.mean(l, axis=1)
        return res

    def _outforecast_draws(self, n_ahead: int, test, expanding: bool, stable, sparse, sv):
        """``McmcOutforecastRun::returnForecast`` (forecaster.h:575-807, 851-875, 920-956): window 0 reuses the fit, windows
        >= 1 are refitted, every (window, chain) in one GPU batch.  Returns (per window: the last-step draws of all chains,
        [chain * draws, k]; per window: the chains' average log-predictive likelihoods, or None)."""
        test = _check_np(test)
        sparse, ci_lev = self._ci_level(sparse)
        k = self.n_features_in_
        n_hor = test.shape[0] - n_ahead + 1
        if n_hor < 1:
            raise ValueError("'test' needs at least 'n_ahead' rows")
        tot = np.asfortranarray(np.vstack([self.y_, test]))
        x_tot, y_tot = self._design_of(tot)
        n0 = self.response_.shape[0]
        if expanding:
            row0 = [0] * n_hor
            nrow = [n0 + w for w in range(n_hor)]
        else:
            row0 = list(range(n_hor))
            nrow = [n0] * n_hor
        seeds = self._seeds(self.chains_, n_hor).T  # _bayes.py: reshape(chains, -1).T
        seed_fc = self._seeds(self.chains_)
        valid = test[n_ahead] if test.shape[0] > n_ahead else None  # y_test.row(step), forecaster.h:802
        C_ = self.chains_
        def per_unit(units, lasts, seeds_u, base):
            """stable=True: the stability filter (subsetStable, config.h:884-914, 1003-1034: companion-matrix eigenvalues
            per draw, decided on the device by bvhar_cuda_is_stable) may keep a different number of draws per unit -> one forecaster call per unit."""
            outs, lps = [], []
            for i, (rec, lo, sd) in enumerate(zip(units, lasts, seeds_u)):
                dens, lp = forecast_density([rec], [lo], n_ahead, self.lag_, self.fit_intercept, self._is_sv, [sd],
                                            har_trans=self._har(), sv=sv, valid_vec=valid, stable=True, sparse=sparse,
                                            device=self._device, unit_id_base=base + i, level=ci_lev)
                outs.append(dens[0][-1].reshape(-1, k))
                lps.append(float(lp.mean()) if lp is not None else None)
            return outs, lps

        # window 0 reuses the existing fit (forecaster.h:799-801)
        lpls = [] if valid is not None else None
        if stable:
            o0, l0 = per_unit(list(self._chain_records()), [tot[: self.lag_ + nrow[0]]] * C_, seed_fc, 0)
            per_window = [np.concatenate(o0, axis=0)]
            if lpls is not None:
                lpls.append(np.asarray(l0))
        else:
            dens0, lpl0 = forecast_density(list(self._chain_records()), [tot[: self.lag_ + nrow[0]]] * C_, n_ahead, self.lag_,
                                           self.fit_intercept, self._is_sv, seed_fc, har_trans=self._har(), sv=sv,
                                           valid_vec=valid, sparse=sparse, device=self._device, level=ci_lev)
            per_window = [np.concatenate([d[-1].reshape(-1, k) for d in dens0], axis=0)]  # last step only (forecaster.h:803)
            if lpls is not None:
                lpls.append(lpl0.mean(axis=1))
        if n_hor > 1 and stable:
            # the draws come back to the host for the filter, then every (window, chain) is forecast on its own
            packed = self._pack(x_tot, y_tot, seeds[1:], win_row0=row0[1:], win_nrow=nrow[1:],
                                record_mask=REC_COEF | REC_SPARSE | REC_LVOL_LAST,
                                lvol_from_init=self._is_sv and expanding, unit_id_base=C_)
            fitw = CudaFit(packed)
            try:
                fitw.run(None)
            except _abi.BvharCudaError as e:
                if e.status != 3:
                    raise
                warnings.warn(str(e))
            bulk = {nm: fitw.record_all(nm, pinned=False) for nm in fitw.record_names()}
            fitw.close()
            for w in range(1, n_hor):
                units = [{nm: arr[(w - 1) * C_ + c] for nm, arr in bulk.items()} for c in range(C_)]
                ow, lw = per_unit(units, [tot[: self.lag_ + row0[w] + nrow[w]]] * C_, seed_fc, C_ + (w - 1) * C_)
                per_window.append(np.concatenate(ow, axis=0))
                if lpls is not None:
                    lpls.append(np.asarray(lw))
        elif n_hor > 1:
            # windows >= 1: refit + forecast of every (window, chain) in ONE native call, the draws stay on the device
            packed = self._pack(x_tot, y_tot, seeds[1:], win_row0=row0[1:], win_nrow=nrow[1:], record_mask=0,
                                lvol_from_init=self._is_sv and expanding, unit_id_base=C_)
            fc, lp = outforecast(packed, tot, n_ahead, self.lag_, seed_fc, har_trans=self._har(), sv=sv, valid_vec=valid,
                                 sparse=sparse, level=ci_lev)
            per_window += [fc[w].reshape(-1, k) for w in range(n_hor - 1)]
            if lpls is not None:
                lpls += [lp[w] for w in range(n_hor - 1)]
        return per_window, lpls

    def roll_forecast(self, n_ahead: int, test, level=.05, stable=False, sparse=False, med=False, sv=True):
        return self._outforecast(n_ahead, test, False, level, stable, sparse, med, sv)

    def expand_forecast(self, n_ahead: int, test, level=.05, stable=False, sparse=False, med=False, sv=True):
        return self._outforecast(n_ahead, test, True, level, stable, sparse, med, sv)


class VarBayes(_AutoregBayes):
    """Bayesian VAR(lag) — arguments as python/src/bvhar/model/_bayes.py:229-246."""

    def __init__(self, data, lag=1, n_chain=1, n_iter=1000, n_burn=None, n_thin=1, bayes_config=None, cov_config=None,
                 intercept_config=None, fit_intercept=True, minnesota=True, ggl=True, verbose=False, n_thread=1, seed=None,
                 device: int = 0, record_mask: Optional[int] = None):
        super().__init__(data, lag, lag, n_chain, n_iter, n_burn, n_thin, bayes_config, cov_config, intercept_config,
                         fit_intercept, "short" if minnesota else "no", ggl, verbose, seed, device, record_mask)
        self.design_ = build_design(self.y_, lag, fit_intercept)
        self.response_ = build_response(self.y_, lag, lag + 1)
        if minnesota:  # _bayes.py:309-310
            self._own_id = np.arange(2, 2 * self.p_, 2, dtype=np.int32)
            self._cross_id = np.arange(1, 2 * self.p_, 2, dtype=np.int32)
            if self._own_id.size == 0:
                self._own_id = np.array([2], dtype=np.int32)
        else:
            self._own_id = np.array([2], dtype=np.int32)
            self._cross_id = np.array([2], dtype=np.int32)
        self._validate()
        self.thread_ = n_thread
        if self.thread_ > n_chain:  # _bayes.py:300-301 (threads have no meaning on the GPU; the message is kept)
            warnings.warn(f"'n_thread = {self.thread_} > 'n_chain' = {n_chain}' will not use every thread. Specify as 'n_thread <= 'n_chain'.")
        self._make_inits()

    def _design_of(self, y):
        return build_design(y, self.lag_, self.fit_intercept), build_response(y, self.lag_, self.lag_ + 1)


class VharBayes(_AutoregBayes):
    """Bayesian VHAR(week, month) — arguments as python/src/bvhar/model/_bayes.py:716-735."""

    def __init__(self, data, week=5, month=22, n_chain=1, n_iter=1000, n_burn=None, n_thin=1, bayes_config=None,
                 cov_config=None, intercept_config=None, fit_intercept=True, minnesota="longrun", ggl=True, verbose=False,
                 n_thread=1, seed=None, device: int = 0, record_mask: Optional[int] = None):
        super().__init__(data, month, 3, n_chain, n_iter, n_burn, n_thin, bayes_config, cov_config, intercept_config,
                         fit_intercept, minnesota, ggl, verbose, seed, device, record_mask)
        self.week_ = week
        self.month_ = month
        self.design_ = build_vhar_design(self.y_, week, month, fit_intercept)
        self.response_ = build_response(self.y_, month, month + 1)
        if minnesota == "longrun":  # _bayes.py:742-750
            self._own_id = np.array([2, 4, 6], dtype=np.int32)
            self._cross_id = np.array([1, 3, 5], dtype=np.int32)
        elif minnesota == "short":
            self._own_id = np.array([2], dtype=np.int32)
            self._cross_id = np.array([1, 3, 4], dtype=np.int32)
        else:
            self._own_id = np.array([1], dtype=np.int32)
            self._cross_id = np.array([2], dtype=np.int32)
        self._validate()
        self.thread_ = n_thread
        if self.thread_ > n_chain:  # _bayes.py:754-755
            warnings.warn(f"'n_thread = {self.thread_} > 'n_chain' = {n_chain}' will not use every thread. Specify as 'n_thread <= 'n_chain'.")
        self._make_inits()

    def _coef_name(self) -> str:
        return "phi"

    def _har(self):
        return build_vhar(self.n_features_in_, self.week_, self.month_, self.fit_intercept)

    def _design_of(self, y):
        return build_vhar_design(y, self.week_, self.month_, self.fit_intercept), build_response(y, self.month_, self.month_ + 1)

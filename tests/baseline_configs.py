"""The five BASELINE.json configurations at their exact shapes (SURVEY.md section 8 size table), with a bounded number
of chains / windows so that the CPU oracle sweeps them in seconds.  Built through the public model classes, i.e. the
packing, group tables, inits and seeds are the ones `var_bayes()` / `vhar_bayes()` / `forecast_roll()` users get.

  C1  VAR(5)  SSVS       LDLT  etf_vix 905 x 9    k=9   d=46  n=900
  C2  VHAR    Horseshoe  SV    etf_vix            k=9   d=28  n=883
  C3  VHAR    DL         SV    synthetic k=20     k=20  d=61  n=1000
  C4  VAR(4)  NG         SV    synthetic k=50     k=50  d=201 n=400   rolling window table
  C5  VHAR    GDP        LDLT  synthetic k=100    k=100 d=301 n=1000
"""
from __future__ import annotations

from bvhar_b200 import _spec as sp
from bvhar_b200 import datasets
from bvhar_b200._abi import REC_ALL
from bvhar_b200.model import VarBayes, VharBayes

SHAPES = {"C1": (9, 46, 900), "C2": (9, 28, 883), "C3": (20, 61, 1000), "C4": (50, 201, 400), "C5": (100, 301, 1000)}
PRIOR = {"C1": "SSVS", "C2": "Horseshoe", "C3": "DL", "C4": "NG", "C5": "GDP"}
SV = {"C1": False, "C2": True, "C3": True, "C4": True, "C5": False}


def packed(cfg: str, chains: int = 2, windows: int = 3, iters: int = 3, burn: int = 0, seed: int = 1, record_mask: int = REC_ALL,
           prior_cfg=None):
    """PackedFit of configuration `cfg` (chains per window; `windows` only matters for C4)."""
    if cfg == "C1":
        m = VarBayes(datasets.load_vix(), 5, chains, iters, burn, 1, prior_cfg or sp.SsvsConfig(), sp.LdltConfig(), seed=seed)
    elif cfg == "C2":
        m = VharBayes(datasets.load_vix(), 5, 22, chains, iters, burn, 1, prior_cfg or sp.HorseshoeConfig(), sp.SvConfig(), seed=seed)
    elif cfg == "C3":
        m = VharBayes(datasets.simulate_vhar_sv(1022, 20), 5, 22, chains, iters, burn, 1, prior_cfg or sp.DlConfig(), sp.SvConfig(), seed=seed)
    elif cfg == "C4":
        y = datasets.simulate_var_sv(404 + windows, 50, 4)
        m = VarBayes(y[:404], 4, chains, iters, burn, 1, prior_cfg or sp.NgConfig(), sp.SvConfig(), seed=seed)
        x_tot, y_tot = m._design_of(y)
        n0 = m.response_.shape[0]
        return m._pack(x_tot, y_tot, m._seeds(windows, chains), win_row0=list(range(1, windows + 1)), win_nrow=[n0] * windows,
                       record_mask=record_mask, unit_id_base=chains)
    elif cfg == "C5":
        m = VharBayes(datasets.simulate_vhar_sv(1022, 100), 5, 22, chains, iters, burn, 1, prior_cfg or sp.GdpConfig(), sp.LdltConfig(), seed=seed)
    else:
        raise KeyError(cfg)
    return m._pack(m.design_, m.response_, m._seeds(1, chains), record_mask=record_mask)

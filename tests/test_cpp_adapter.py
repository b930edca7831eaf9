"""The C++ host adapter (include/bvhar_b200/mcmc_run.hpp, the mirror of bvhar::McmcRun, triangular.h:1192-1233):
compiles with g++ against the C-ABI, refuses loudly without a GPU (CPU test), and reproduces the Python adapter's
records bit for bit on a GPU (gpu test)."""
import os
import struct
import subprocess

import numpy as np
import pytest

import problems as pr
from bvhar_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "_build", "adapter_main")


def _build():
    src = os.path.join(ROOT, "tests", "cpp", "adapter_main.cpp")
    hdr = os.path.join(ROOT, "include", "bvhar_b200", "mcmc_run.hpp")
    hdr2 = os.path.join(ROOT, "include", "bvhar_b200", "forecast_run.hpp")
    if os.path.exists(EXE) and os.path.getmtime(EXE) >= max(os.path.getmtime(src), os.path.getmtime(hdr), os.path.getmtime(hdr2),
                                                             os.path.getmtime(_abi.lib_path())):
        return
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    libdir = os.path.dirname(_abi.lib_path())
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), src, "-o", EXE,
                    "-L", libdir, "-lbvhar_b200", f"-Wl,-rpath,{libdir}"], check=True)


def _write(path, arrays):
    with open(path, "wb") as f:
        for name, a in arrays.items():
            a = np.asfortranarray(np.atleast_2d(np.asarray(a, dtype=np.float64)))
            if a.shape[0] == 1 and a.size > 1:
                a = np.asfortranarray(a.T)
            nb = name.encode()
            f.write(struct.pack("<i", len(nb))); f.write(nb)
            f.write(struct.pack("<qq", a.shape[0], a.shape[1])); f.write(a.tobytes(order="F"))


def _read(path):
    out = {}
    with open(path, "rb") as f:
        while True:
            h = f.read(4)
            if len(h) < 4:
                break
            (nl,) = struct.unpack("<i", h)
            name = f.read(nl).decode()
            r, c = struct.unpack("<qq", f.read(16))
            out[name] = np.frombuffer(f.read(8 * r * c), dtype=np.float64).reshape(r, c, order="F")
    return out


def _problem_file(tmp_path, prior, sv, forecast=None, **kw):
    kwp, meta = pr.make_problem(prior=prior, sv=sv, **kw)
    arrays = {"meta": [kwp["num_chains"], kwp["num_iter"], kwp["num_burn"], kwp["thin"], kwp["prior_type"],
                       int(kwp["include_mean"]), int(sv), int(kwp["is_group"])],
              "x": kwp["x"], "y": kwp["y"], "grp_id": kwp["grp_id"], "own_id": kwp["own_id"], "cross_id": kwp["cross_id"],
              "grp_mat": np.asarray(kwp["grp_mat"]).reshape(-1, order="F"), "seed_chain": np.asarray(kwp["seed_chain"]).reshape(-1)}
    for pre, dic in (("cov.", kwp["param_cov"]), ("prior.", kwp["param_prior"]), ("icpt.", kwp["param_intercept"])):
        for k, v in dic.items():
            v = np.asarray(v, dtype=np.float64)
            arrays[pre + k] = v if v.ndim == 2 else v.reshape(-1)
    for c, ini in enumerate(kwp["param_init"]):
        for k, v in ini.items():
            v = np.asarray(v, dtype=np.float64)
            arrays[f"init{c}.{k}"] = v if v.ndim == 2 else v.reshape(-1)
    if forecast is not None:
        arrays["forecast"] = np.asarray(forecast, dtype=np.float64)
    path = os.path.join(tmp_path, "problem.bin")
    _write(path, arrays)
    return path, kwp


def test_adapter_compiles_and_refuses_without_gpu(tmp_path):
    _build()
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    path, _ = _problem_file(str(tmp_path), "Horseshoe", True, chains=2, iters=4, burn=1)
    r = subprocess.run([EXE, path, os.path.join(str(tmp_path), "out.bin")], capture_output=True, text=True)
    assert r.returncode == 3 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("prior,sv", [("Horseshoe", True), ("Minnesota", False), ("HMN", True), ("SSVS", False), ("NG", True),
                                      ("DL", False), ("GDP", True)])
def test_cpp_adapter_matches_python_adapter(tmp_path, prior, sv):
    _build()
    path, kwp = _problem_file(str(tmp_path), prior, sv, chains=2, iters=5, burn=1)
    out = os.path.join(str(tmp_path), "out.bin")
    r = subprocess.run([EXE, path, out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = _read(out)
    from bvhar_b200._engine import CudaFit
    g = CudaFit(_abi.PackedFit(**kwp)); g.run()
    for c in range(2):
        for nm in g.record_names():
            if prior in ("Minnesota", "HMN"):  # the prior moments are computed by each adapter: equal to rounding
                np.testing.assert_allclose(got[f"c{c}.{nm}"], g.record(nm, 0, c), rtol=1e-8, atol=1e-12, equal_nan=True)
            else:
                assert np.array_equal(got[f"c{c}.{nm}"], g.record(nm, 0, c), equal_nan=True), (prior, nm)


@pytest.mark.gpu
@pytest.mark.parametrize("vhar,sv,stable", [(False, False, False), (False, True, True), (True, True, False), (True, False, True)])
def test_cpp_forecast_adapter_matches_python_adapter(tmp_path, vhar, sv, stable):
    """bvhar_b200::McmcForecastRun (include/bvhar_b200/forecast_run.hpp, the mirror of forecaster.h:494-568) on the records
    of the C++ McmcRun equals the Python forecast_density on the Python adapter's records, bit for bit."""
    _build()
    from bvhar_b200._design import build_vhar
    from bvhar_b200._engine import CudaFit, forecast_density
    step = 3
    kw = dict(chains=2, iters=8, burn=2, dim=3, T=60, vhar=vhar, week=2, month=4)
    fc = [2, 4, step, int(stable)] if vhar else [2, 0, step, int(stable)]
    path, kwp = _problem_file(str(tmp_path), "Horseshoe", sv, forecast=fc, **kw)
    out = os.path.join(str(tmp_path), "out.bin")
    r = subprocess.run([EXE, path, out], capture_output=True, text=True)
    if r.returncode == 3 and ("same number of draws" in r.stderr or "No stable MCMC draws" in r.stderr):
        pytest.skip(r.stderr.strip())
    assert r.returncode == 0, r.stderr
    got = _read(out)
    g = CudaFit(_abi.PackedFit(**kwp)); g.run()
    lag = 4 if vhar else 2
    har = build_vhar(3, 2, 4, True) if vhar else None
    try:
        dens, _ = forecast_density([g.records(0, c) for c in range(2)], [np.asarray(kwp["y"])] * 2, step, lag, True, sv,
                                   np.asarray(kwp["seed_chain"]).reshape(-1), har_trans=har, stable=stable)
    except ValueError as e:  # the stability filter kept different numbers of draws per chain: both adapters refuse
        pytest.skip(str(e))
    for c in range(2):
        assert np.array_equal(got[f"f{c}"], dens[c])


@pytest.mark.gpu
@pytest.mark.parametrize("vhar,sv,expanding,stable,level", [(False, True, False, False, 0.0), (False, False, True, False, 0.0),
                                                            (True, True, True, False, 0.0), (True, False, False, False, 0.3),
                                                            (False, True, False, True, 0.0)])
def test_cpp_outforecast_runners_match_python_model(tmp_path, vhar, sv, expanding, stable, level):
    """bvhar_b200::McmcVarforecastRun / McmcVharforecastRun<Roll | Expand, ...> (include/bvhar_b200/outforecast_run.hpp, the
    mirrors of forecaster.h:966-1121) against `model.py::_outforecast_draws` on the same fit, seeds and test rows: the
    forecast draws of every (window, chain) and the lpl table, bit for bit."""
    _build()
    from bvhar_b200 import _spec as sp
    from bvhar_b200.model import VarBayes, VharBayes
    y = pr.simulate_series(64, 3, 4242)
    train, test = y[:58], y[58:]
    step, chains, iters, burn = 2, 2, 12, 4
    cov = sp.SvConfig() if sv else sp.LdltConfig()
    if vhar:
        m = VharBayes(train, 2, 4, chains, iters, burn, 1, sp.HorseshoeConfig(), cov, seed=5)
    else:
        m = VarBayes(train, 2, chains, iters, burn, 1, sp.HorseshoeConfig(), cov, seed=5)
    m.fit()
    n_hor = test.shape[0] - step + 1
    fixed = {(chains, n_hor): np.arange(100, 100 + chains * n_hor).reshape(chains, n_hor), (chains,): np.arange(7, 7 + chains)}
    m._seeds = lambda *shape: fixed[shape]
    sparse_arg = level if level > 0 else False
    try:
        per_window, lpls = m._outforecast_draws(step, test, expanding, stable, sparse_arg, True)
    except ValueError as e:
        pytest.skip(str(e))
    # the same job through the C++ runner: the model's own constructor arguments and fit records
    arrays = {"meta": [chains, iters, burn, 1, int(m._prior_type), 1, int(sv), 1],
              "x": m.design_, "y": m.response_, "y_train": train, "y_test": test,
              "grp_id": m._group_id, "own_id": m._own_id, "cross_id": m._cross_id, "grp_mat": np.asarray(m.group_).reshape(-1, order="F"),
              "seed_chain": np.arange(1, chains + 1), "seed_window_chain": fixed[(chains, n_hor)].T.reshape(-1), "seed_forecast": fixed[(chains,)],
              "outforecast": [int(vhar), int(expanding), 2, 4 if vhar else 0, step, int(stable), level, 1, 0]}
    for pre, dic in (("cov.", m.cov_spec_.to_dict()), ("prior.", m.spec_.to_dict()), ("icpt.", m.intercept_spec_.to_dict())):
        for k2, v in dic.items():
            v = np.asarray(v, dtype=np.float64)
            arrays[pre + k2] = v if v.ndim == 2 else v.reshape(-1)
    for c, ini in enumerate(m.init_):
        for k2, v in ini.items():
            v = np.asarray(v, dtype=np.float64)
            arrays[f"init{c}.{k2}"] = v if v.ndim == 2 else v.reshape(-1)
    for c in range(chains):
        for nm, arr in m.param_.items():
            arrays[f"rec{c}.{nm}"] = np.asarray(arr[c])
    path = os.path.join(str(tmp_path), "problem.bin")
    _write(path, arrays)
    out = os.path.join(str(tmp_path), "out.bin")
    r = subprocess.run([EXE, path, out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = _read(out)
    k = 3
    for w in range(n_hor):
        mine = np.concatenate([got[f"o{w}.{c}"].reshape(-1, k) for c in range(chains)], axis=0)
        assert np.array_equal(mine, per_window[w]), (w, np.abs(mine - per_window[w]).max() if mine.shape == per_window[w].shape else (mine.shape, per_window[w].shape))
    assert np.array_equal(got["lpl"], np.stack(lpls))

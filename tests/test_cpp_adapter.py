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

"""The drop-in boundary: include/bvhar_cuda.h <-> libbvhar_b200.so <-> the ctypes mirror.  No GPU needed
(no compute call is made); on a box without a device the engine must refuse loudly, never fall back."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from bvhar_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "bvhar_cuda.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bvhar_cuda_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_entry_points():
    syms = _declared_symbols()
    for must in ("bvhar_cuda_fit_create", "bvhar_cuda_fit_run", "bvhar_cuda_fit_record", "bvhar_cuda_fit_destroy",
                 "bvhar_cuda_forecast_density", "bvhar_cuda_last_error", "bvhar_cuda_outforecast_run"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(_abi.lib_path())
    missing = [s for s in _declared_symbols() if not hasattr(lib, s)]
    assert not missing, f"declared in include/bvhar_cuda.h but not exported: {missing}"


def test_abi_version_and_struct_layout():
    L = _abi.load_library()  # raises if the version or any sizeof() differs
    assert L.bvhar_cuda_abi_version() == _abi.ABI_VERSION
    m = re.search(r"#define\s+BVHAR_CUDA_ABI_VERSION\s+(\d+)", open(HEADER).read())
    assert int(m.group(1)) == _abi.ABI_VERSION


def test_enums_match_header():
    src = open(HEADER).read()
    for name, val in (("BVHAR_PRIOR_MINNESOTA", 1), ("BVHAR_PRIOR_SSVS", 2), ("BVHAR_PRIOR_HORSESHOE", 3),
                      ("BVHAR_PRIOR_HIERMINN", 4), ("BVHAR_PRIOR_NG", 5), ("BVHAR_PRIOR_DL", 6), ("BVHAR_PRIOR_GDP", 7),
                      ("BVHAR_COV_LDLT", 0), ("BVHAR_COV_SV", 1)):
        assert re.search(rf"{name}\s*=\s*{val}\b", src), name


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the refusal path is exercised on CPU-only boxes")
    import problems as pr
    pk, _ = pr.pack(prior="Minnesota", sv=False)
    L = _abi.load_library()
    h = C.c_void_p()
    rc = L.bvhar_cuda_fit_create(C.byref(pk.desc), C.byref(h))
    assert rc == 2, "fit_create must fail with BVHAR_ERR_CUDA when no device is usable"
    assert b"no CPU fallback" in L.bvhar_cuda_last_error()
    out = np.zeros(4)
    a = np.ones((2, 2))
    rc = L.bvhar_cuda_dgemm_tn(0, 2, 2, 2, a.ctypes.data_as(_abi.c_dp), 2, a.ctypes.data_as(_abi.c_dp), 2,
                               out.ctypes.data_as(_abi.c_dp), 2)
    assert rc == 2


def test_product_package_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing in the product may import, include, link or dlopen it."""
    pkg = os.path.join(ROOT, "bvhar_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if not f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                continue
            for line in open(os.path.join(dirpath, f), errors="ignore"):
                s = line.strip()
                if s.startswith(("import ", "from ", "#include")) or "CDLL" in s or "dlopen" in s or "-l" in s.split():
                    assert "oracle" not in s.lower(), f"{f}: {s}"


def test_descriptor_validation_is_host_side():
    import problems as pr
    kw, _ = pr.make_problem(prior="SSVS", sv=True)
    bad = dict(kw)
    bad["seed_chain"] = np.arange(5)
    with pytest.raises(ValueError):
        _abi.PackedFit(**bad)
    bad = dict(kw)
    bad["grp_mat"] = np.ones((3, 3), dtype=np.int32)
    with pytest.raises(ValueError):
        _abi.PackedFit(**bad)
    bad = dict(kw)
    bad["param_init"] = kw["param_init"][:1]
    with pytest.raises(ValueError):
        _abi.PackedFit(**bad)


@pytest.mark.parametrize("prior", ["Minnesota", "HMN", "SSVS", "Horseshoe", "NG", "DL", "GDP"])
@pytest.mark.parametrize("sv", [True, False])
def test_chain_init_table_matches_the_inputs(prior, sv):
    """PackedFit fills bvhar_chain_init column by column through numpy views of the struct array (thousands of chains): every
    pointer field must address exactly the array the chain passed (no copy for float64 vectors, same values otherwise) and
    every scalar field must carry the chain's value - read back here through the ctypes field descriptors."""
    import problems as pr
    kw, _ = pr.make_problem(prior=prior, sv=sv, dim=3, T=40, chains=5)
    pk = _abi.PackedFit(**kw)
    inits = C.cast(pk.desc.init, C.POINTER(_abi.ChainInit))
    ptr_keys = {"init_coef": "init_coef", "init_contem": "init_contem", "init_diag": "init_diag", "lvol_init": "lvol_init",
                "lvol": "lvol", "lvol_sig": "lvol_sig", "coef_dummy": "init_coef_dummy", "coef_mixture": "coef_mixture",
                "chol_mixture": "chol_mixture", "coef_slab": "coef_slab", "contem_slab": "contem_slab",
                "local_sparsity": "local_sparsity", "group_sparsity": "group_sparsity",
                "contem_local_sparsity": "contem_local_sparsity", "local_shape": "local_shape", "group_rate": "group_rate",
                "contem_rate": "contem_rate"}
    scalar_keys = ["own_lambda", "cross_lambda", "contem_lambda", "coef_spike_scl", "chol_spike_scl", "global_sparsity",
                   "contem_global_sparsity", "contem_shape", "gamma_shape", "gamma_rate", "contem_gamma_shape", "contem_gamma_rate"]
    used = 0
    for c, ini in enumerate(kw["param_init"]):
        I = inits[c]
        for field, key in ptr_keys.items():
            p = getattr(I, field)
            if key not in ini:
                continue
            if not p:  # a key the prior / covariance model does not read stays NULL
                continue
            a = np.asarray(ini[key], dtype=np.float64)
            got = np.ctypeslib.as_array(p, shape=(a.size,))
            np.testing.assert_array_equal(got, a.reshape(-1, order="F"))
            if isinstance(ini[key], np.ndarray) and ini[key].dtype == np.float64 and ini[key].ndim == 1:
                assert C.addressof(p.contents) == ini[key].__array_interface__["data"][0]  # passed through, not copied
            used += 1
        for field in scalar_keys:
            if field in ini and getattr(I, field) != 0.0:
                assert getattr(I, field) == float(np.asarray(ini[field], dtype=np.float64).reshape(-1)[0])
                used += 1
    assert used >= 2 * 5


def test_chain_init_pointers_for_awkward_arrays():
    """The pointer columns are read through the buffer protocol (writable, C-contiguous buffers) with a fallback for the
    rest: read-only vectors, row-major or column-major matrices, Python lists and float32 inputs must all end up as
    pointers to the right values."""
    import problems as pr
    kw, _ = pr.make_problem(prior="Horseshoe", sv=True, dim=3, T=40, chains=4)
    inits_in = kw["param_init"]
    ro = np.array(inits_in[0]["lvol_init"], dtype=np.float64)
    ro.setflags(write=False)
    inits_in[0]["lvol_init"] = ro                                                        # read-only: fallback path
    inits_in[1]["init_coef"] = np.ascontiguousarray(inits_in[1]["init_coef"])           # row-major matrix: copied to column-major
    inits_in[2]["init_coef"] = np.asfortranarray(inits_in[2]["init_coef"])              # column-major: passed through
    inits_in[3]["init_contem"] = [float(v) for v in np.asarray(inits_in[3]["init_contem"]).reshape(-1)]  # list
    inits_in[3]["local_sparsity"] = np.asarray(inits_in[3]["local_sparsity"], dtype=np.float32)         # float32
    pk = _abi.PackedFit(**kw)
    inits = C.cast(pk.desc.init, C.POINTER(_abi.ChainInit))
    for c, ini in enumerate(inits_in):
        for field in ("lvol_init", "init_coef", "init_contem", "local_sparsity", "lvol"):
            a = np.asarray(ini[field], dtype=np.float64)
            got = np.ctypeslib.as_array(getattr(inits[c], field), shape=(a.size,))
            np.testing.assert_array_equal(got, a.reshape(-1, order="F"), err_msg=f"chain {c} field {field}")
    assert C.addressof(inits[0].lvol_init.contents) == ro.__array_interface__["data"][0]
    assert C.addressof(inits[2].init_coef.contents) == inits_in[2]["init_coef"].__array_interface__["data"][0]

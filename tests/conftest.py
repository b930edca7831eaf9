import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `pytest -m gpu`)")


def pytest_collection_modifyitems(config, items):
    """`-m gpu` tests need a CUDA device AND the built library: without either they are skipped, not failed
    (the product itself still fails loudly - see tests/test_abi.py)."""
    import glob
    have_lib = bool(glob.glob(os.path.join(ROOT, "bvhar_b200", "csrc", "libbvhar_b200.so")))
    have_gpu = False
    if have_lib:
        import ctypes
        try:
            lib = ctypes.CDLL(os.path.join(ROOT, "bvhar_b200", "csrc", "libbvhar_b200.so"))
            n = ctypes.c_int(0)
            have_gpu = lib.bvhar_cuda_device_count(ctypes.byref(n)) == 0 and n.value > 0
        except OSError:
            have_lib = False
    if have_gpu and have_lib:
        return
    why = "no CUDA device" if not have_gpu else "libbvhar_b200.so not built"
    skip = pytest.mark.skip(reason=f"gpu test: {why}")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _build_oracle():
    """Compile the CPU oracle once per session (g++, a few seconds)."""
    import oracle_binding
    oracle_binding.build()

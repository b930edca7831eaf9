"""The oracle against its committed golden trajectories (CPU): guards the checker itself against drift."""
import numpy as np
import pytest
from numpy.testing import assert_allclose

import golden
import oracle_binding as ob
import problems as pr


@pytest.mark.parametrize("name", sorted(golden.CASES))
def test_oracle_reproduces_golden(name):
    case = golden.load(name)
    kw, meta = pr.make_problem(**case["spec"])
    assert_allclose(np.abs(meta["x"]).sum(), case["x_checksum"], rtol=1e-13), "the seeded test problem itself changed"
    from bvhar_b200._abi import PackedFit
    for faithful in (False, True):
        f = ob.OracleFit(PackedFit(**kw), faithful=faithful); f.run()
        for nm, exp in case["records"].items():
            got = np.stack([f.record(nm, w, c) for w in range(case["windows"]) for c in range(case["chains"])])
            assert_allclose(got, exp, rtol=1e-7 if faithful else 1e-10, atol=1e-10, equal_nan=True, err_msg=f"{name} {nm} faithful={faithful}")

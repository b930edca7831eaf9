"""Regenerate tests/golden/*.npz from the CPU oracle:  python tests/golden/make_golden.py
(the reference cannot be imported or built in this image - no R/Rcpp/Eigen/Boost - so the fixtures are
outputs of the oracle restatement, with the generating problem spec in tests/golden.py)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import golden  # noqa: E402
import oracle_binding as ob  # noqa: E402
import problems as pr  # noqa: E402
from bvhar_b200._abi import PackedFit  # noqa: E402

for name, spec in golden.CASES.items():
    kw, meta = pr.make_problem(**spec)
    f = ob.OracleFit(PackedFit(**kw)); f.run()
    W = len(spec.get("windows") or [0]); Cn = spec["chains"]
    out = {"x_checksum": np.float64(np.abs(meta["x"]).sum())}
    names = [n for n in f.record_names() if n in golden.RECORDS or n in ("d_record", "sigh_record", "h0_record", "lambda_record", "gamma_record")]
    for nm in names:
        out["rec_" + nm] = np.stack([f.record(nm, w, c) for w in range(W) for c in range(Cn)])
    np.savez_compressed(golden.path(name), **out)
    print(name, {k: v.shape for k, v in out.items() if k != "x_checksum"})

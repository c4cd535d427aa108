"""Oracle regression fixture (oracle-generated, NOT reference-pinned -- tests/golden/README.md)."""
import importlib.util
import os

import numpy as np

from conftest import ROOT, rel_err


def test_oracle_matches_its_committed_fixture():
    spec = importlib.util.spec_from_file_location("mk", os.path.join(ROOT, "tests", "golden", "make_oracle_regression.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    got = mk.compute()
    ref = np.load(os.path.join(ROOT, "tests", "golden", "oracle_regression_f64.npz"))
    for k in ("scaling", "b", "inc", "block7", "jl_col_scale7", "cams", "lms", "lm_costs"):
        assert rel_err(got[k], ref[k]) < 1e-10, k   # compilers may contract FMAs differently; far above that it is a real change
    for k in ("cost0", "cost1", "l_diff"):
        assert abs(got[k] - float(ref[k])) <= 1e-10 * abs(float(ref[k])), k
    assert int(got["cg_iterations"]) == int(ref["cg_iterations"])
    assert np.all(np.abs(got["lm_cg"] - ref["lm_cg"]) <= 1)

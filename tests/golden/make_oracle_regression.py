#!/usr/bin/env python
"""Regenerates tests/golden/oracle_regression_f64.npz from the CPU oracle (NOT from the reference, see README.md)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_py as orc  # noqa: E402
from rootba_b200.synthetic import synth_bal  # noqa: E402


def compute():
    a = synth_bal(10, 120, 3.6, seed=2026)
    o = orc.Oracle(a, np.float64, orc.default_options(num_threads=1))
    e0 = o.compute_error()
    assert o.linearize()
    scaling = o.get_scaling()
    inc, dbg = o.solve(1e-4, want_debug=True)
    blk, lm_idx, res_idx, jls = o.get_block(7)
    l_diff = o.apply(inc.copy())
    cams, lms = o.get_state()
    e1 = o.compute_error()
    o2 = orc.Oracle(a, np.float64, orc.default_options(num_threads=1, max_num_iterations=4))
    rows, term = o2.optimize()
    return dict(cost0=e0["all"]["error"], scaling=scaling, b=dbg["b"], inc=inc, cg_iterations=dbg["cg_iterations"], block7=blk,
                jl_col_scale7=jls, l_diff=l_diff, cams=cams, lms=lms, cost1=e1["all"]["error"],
                lm_costs=np.array([r["cost"] for r in rows]), lm_cg=np.array([r["cg_iterations"] for r in rows]))


if __name__ == "__main__":
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_regression_f64.npz"), **compute())
    print("written")

#!/usr/bin/env python
"""Tiny driver for ncu: linearize + one solve + a few operator launches on a BASELINE stand-in.
   python scripts/prof_operator.py [workload] [dense|implicit] [f32|f64]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rootba_b200 as rb
from rootba_b200.synthetic import synth_config

wl = sys.argv[1] if len(sys.argv) > 1 else "ladybug-1723"
form = (sys.argv[2] if len(sys.argv) > 2 else "dense").upper()
dtype = np.float64 if (len(sys.argv) > 3 and sys.argv[3] == "f64") else np.float32
bp = rb.BalProblem.from_arrays(synth_config(wl), dtype)
so = rb.SolverOptions(operator_form=form, max_linear_solver_iterations=3)
lin = rb.LinearizorQR.create(bp, so)
lin.compute_error(); lin.linearize(); lin.solve(1e-4)
print("operator us:", lin.time_matvec(5) * 1e6)

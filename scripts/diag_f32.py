#!/usr/bin/env python
"""Diagnostic (GPU box): where do the float32 GPU and float32 oracle LM trajectories part?
Prints, per LM iteration, cost / lambda / cg iterations / accept for GPU-f32, oracle-f32 and oracle-f64, and the
single-stage quantities (b, SCHUR_JACOBI inverse blocks, H x, inc, l_diff) of GPU-f32 and oracle-f32 against oracle-f64."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import rootba_b200 as rb  # noqa: E402
from oracle import oracle_py as orc  # noqa: E402
from rootba_b200.synthetic import synth_bal  # noqa: E402


def rel(a, b):
    a = np.asarray(a, np.float64).ravel(); b = np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / (np.linalg.norm(a) + np.linalg.norm(b) + 1e-300))


def stage(arrays, lam, tag, form="PANEL"):
    tag = f"{tag}/{form}"
    o64 = orc.Oracle(arrays, np.float64, orc.default_options(num_threads=0))
    o32 = orc.Oracle(arrays, np.float32, orc.default_options(num_threads=0))
    bp = rb.BalProblem.from_arrays(arrays, np.float32)
    lin = rb.LinearizorQR.create(bp, rb.SolverOptions(stage2_form=form))
    lin.linearize(); o64.linearize(); o32.linearize()
    ig = lin.solve(lam); i64, d64 = o64.solve(lam, want_debug=True); i32, d32 = o32.solve(lam, want_debug=True)
    invg, _ = lin.get_preconditioner()
    x = np.random.default_rng(3).uniform(-1, 1, 9 * lin.nc)
    print(f"[{tag} lam={lam}] b: gpu {rel(lin.get_rhs(), d64['b']):.2e} orc32 {rel(d32['b'], d64['b']):.2e} | "
          f"inv(worst): gpu {max(rel(invg[c], d64['inv_blocks'][c]) for c in range(lin.nc)):.2e} "
          f"orc32 {max(rel(d32['inv_blocks'][c], d64['inv_blocks'][c]) for c in range(lin.nc)):.2e} | "
          f"Hx: gpu {rel(lin.right_multiply(x.astype(np.float32)), o64.right_multiply(x)):.2e} "
          f"orc32 {rel(o32.right_multiply(x.astype(np.float32)), o64.right_multiply(x)):.2e} | "
          f"inc: gpu {rel(ig, i64):.2e} orc32 {rel(i32, i64):.2e} gpu-vs-orc32 {rel(ig, i32):.2e} | "
          f"cg gpu {lin.last_cg.num_iterations} o32 {d32['cg_iterations']} o64 {d64['cg_iterations']}")
    lg, l32, l64 = lin.apply(ig), o32.apply(i32), o64.apply(i64)
    print(f"    l_diff gpu {lg:.8e} o32 {l32:.8e} o64 {l64:.8e}")
    lin.close()


def traj(arrays, tag, iters=8, form="PANEL"):
    tag = f"{tag}/{form}"
    o64 = orc.Oracle(arrays, np.float64, orc.default_options(num_threads=0, max_num_iterations=iters))
    o32 = orc.Oracle(arrays, np.float32, orc.default_options(num_threads=0, max_num_iterations=iters))
    r64, _ = o64.optimize(); r32, _ = o32.optimize()
    bp = rb.BalProblem.from_arrays(arrays, np.float32)
    so = rb.SolverOptions(max_num_iterations=iters, stage2_form=form)
    summ = rb.bundle_adjust_manual(bp, so)
    g = summ["iterations"]
    print(f"[{tag}] rows: gpu {len(g)} o32 {len(r32)} o64 {len(r64)}")
    for k in range(max(len(g), len(r32), len(r64))):
        s = f"  it {k:2d}"
        if k < len(g):
            a = g[k]
            s += f" | gpu cost {a['cost']['all']['error']:.9e} lam {a.get('lam', 0):.3e} cg {a.get('linear_solver_iterations', -1):4d} ok {int(bool(a.get('step_is_successful')))} ld {a.get('l_diff', 0):.6e}"
        for name, r in (("o32", r32), ("o64", r64)):
            if k < len(r):
                b = r[k]
                s += f" | {name} cost {b['cost']:.9e} lam {b['lambda']:.3e} cg {int(b['cg_iterations']):4d} ok {int(b['step_is_successful'])} ld {b['l_diff']:.6e}"
        print(s)


if __name__ == "__main__":
    small = synth_bal(49, 1800, 4.1, seed=38401)
    mixed = synth_bal(150, 1500, 9.0, seed=11, max_track=150)
    for form in ("PANEL", "IDENTITY"):
        for lam in (0.1, 1e-4):
            stage(small, lam, "small", form)
            stage(mixed, lam, "mixed", form)
    for form in ("PANEL", "IDENTITY"):
        traj(small, "small", form=form)
        traj(mixed, "mixed", form=form)

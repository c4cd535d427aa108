"""CPU-side checks of the C-ABI library: it loads without a GPU, exports every symbol the public header
declares, refuses to run without a device (no CPU fallback), and its host-only helpers work."""
import ctypes as C

import numpy as np
import pytest

import rootba_b200 as rb
from rootba_b200 import _lib


def test_library_exports_every_declared_symbol():
    _lib.build()
    L = _lib.lib()
    syms = _lib.declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/rootba_b200.h but not exported"
    assert L.rba_abi_version() == 1


def test_no_cpu_fallback(tiny_problem):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    bp = rb.BalProblem.from_arrays(tiny_problem)
    with pytest.raises(rb.RbaError) as e:
        rb.LinearizorQR.create(bp, rb.SolverOptions())
    assert e.value.code == -2  # RBA_ERR_NO_DEVICE


def test_partition_landmarks(small_problem):
    off = small_problem.lm_off
    n = np.diff(off).astype(np.float64)
    w = n * n + 4 * n
    for nranks in (1, 2, 3, 8):
        b = rb.partition_landmarks(off, nranks)
        assert b[0] == 0 and b[-1] == small_problem.nl and np.all(np.diff(b) >= 0)
        loads = np.array([w[b[r]:b[r + 1]].sum() for r in range(nranks)])
        assert loads.max() <= loads.mean() * 1.05 + w.max()


def test_default_opts_match_reference_defaults():
    o = _lib.SolverOpts()
    _lib.lib().rba_default_solver_opts(C.byref(o))
    # bal/solver_options.hpp: SCHUR_JACOBI, max 500 its, eta 0.1, Householder
    assert (o.preconditioner_type, o.max_linear_solver_iterations, o.eta, o.use_householder_marginalization) == (1, 500, 0.1, 1)
    assert o.residual_reset_period == 10 and o.nranks == 1

"""Host-side mirror of the reference's interface for the QR hot path, on top of the C ABI.

  BalProblem        <-> rootba::BalProblem<Scalar>     (src/rootba/bal/bal_problem.hpp:61-234)
  SolverOptions     <-> rootba::SolverOptions          (src/rootba/bal/solver_options.hpp:46-284)
  LinearizorQR      <-> rootba::LinearizorQR<Scalar>   (src/rootba/solver/linearizor_qr.cpp:52-291)
                        behind rootba::Linearizor      (src/rootba/solver/linearizor.hpp:47-83)

Same names, argument meaning and error behaviour: numerical trouble is reported through return values
(NaN l_diff, non-finite increment), invariant violations raise (the reference CHECK-aborts).
The optimisation state lives on the GPU; `BalProblem.cameras/landmarks` are refreshed on demand.
"""
from __future__ import annotations

import ctypes as C
import os
import dataclasses
import math
import time

import numpy as np

from . import _lib
from ._lib import (CgSummary, LmIteration, LmOpts, LmStepResult, ProblemView, RbaError, ResidualInfo, SolverOpts, StageTimings, WorkloadStats,
                   check, struct_to_dict)


@dataclasses.dataclass
class ResidualOptions:  # bal/bal_residual_options.hpp:44-63
    robust_norm: str = "NONE"      # NONE | HUBER
    huber_parameter: float = 1.0


@dataclasses.dataclass
class SolverOptions:  # bal/solver_options.hpp (QR-relevant subset, reference defaults)
    solver_type: str = "SQUARE_ROOT"              # SQUARE_ROOT | SCHUR_COMPLEMENT | POWER_SCHUR_COMPLEMENT (solver_options.hpp:63-76)
    power_order: int = 20                         # :270
    optimized_cost: str = "ERROR"                 # ERROR | ERROR_VALID | ERROR_VALID_AVG
    max_num_iterations: int = 20
    min_relative_decrease: float = 0.0           # solver_options.hpp:146-148
    initial_trust_region_radius: float = 1e4
    min_trust_region_radius: float = 1e-32
    max_trust_region_radius: float = 1e16
    min_linear_solver_iterations: int = 0
    max_linear_solver_iterations: int = 500
    eta: float = 0.1
    jacobi_scaling_epsilon: float = 0.0
    preconditioner_type: str = "SCHUR_JACOBI"     # JACOBI | SCHUR_JACOBI
    function_tolerance: float = 1e-6
    use_double: bool = True
    use_householder_marginalization: bool = True
    staged_execution: bool = True
    reduction_alg: int = 1
    initial_vee: float = 2.0
    vee_factor: float = 2.0
    residual: ResidualOptions = dataclasses.field(default_factory=ResidualOptions)
    # placement (not in the reference)
    device: int = -1
    rank: int = 0
    nranks: int = 1
    pcg_check_period: int = 4
    operator_form: str = "DENSE"                  # DENSE (reference: Q2 panels) | IMPLICIT (Jp^T Jp - Q1d^T Q1d from records)
    stage2_form: str = "PANEL"                    # PANEL (reference: b and SCHUR_JACOBI blocks from the Q2 panels) | IDENTITY

    def use_projection_validity_check(self) -> bool:  # solver_options.cpp:41-51
        return self.optimized_cost != "ERROR"


class BalProblem:
    """SoA BalProblem: cameras [nc,10] (quat xyzw, t, f,k1,k2), landmarks [nl,3], observations in
    CSR-by-landmark order with ascending camera index."""

    def __init__(self, cams, lms, lm_off, obs_cam, obs_xy, dtype=np.float64):
        self.dtype = np.dtype(dtype)
        # the problem owns its optimisation state (it is updated in place by the solver)
        self.cams = np.array(cams, dtype=self.dtype, order="C", copy=True).reshape(-1, 10)
        self.lms = np.array(lms, dtype=self.dtype, order="C", copy=True).reshape(-1, 3)
        self.lm_off = np.ascontiguousarray(lm_off, dtype=np.int64)
        self.obs_cam = np.ascontiguousarray(obs_cam, dtype=np.int32)
        self.obs_xy = np.ascontiguousarray(obs_xy, dtype=self.dtype).reshape(-1, 2)
        self._cams_backup = self.cams.copy()
        self._lms_backup = self.lms.copy()
        self._linearizor = None

    @classmethod
    def from_arrays(cls, arrays, dtype=np.float64) -> "BalProblem":
        return cls(arrays.cams, arrays.lms, arrays.lm_off, arrays.obs_cam, arrays.obs_xy, dtype)

    @classmethod
    def load_bal(cls, path: str, dtype=np.float64, normalize: bool = True, scale: float = 100.0, num_threads: int = 0,
                 init_depth_threshold: float = 0.0, rotation_sigma: float = 0.0, translation_sigma: float = 0.0,
                 point_sigma: float = 0.0, random_seed: int = 38401) -> "BalProblem":
        """load_normalized_bal_problem (bal/bal_problem.cpp:773-852) through the library's multi-threaded BAL parser
        (rba_bal_load): load + normalise in double, then cast to dtype."""
        L = _lib.lib()
        f = C.c_void_p()
        check(L.rba_bal_load(os.fsencode(path), int(normalize), C.c_double(scale), int(num_threads), C.byref(f)))
        try:
            if rotation_sigma > 0 or translation_sigma > 0 or point_sigma > 0:  # BalProblem::perturb (bal_problem.cpp:507-554, :820-822)
                check(L.rba_bal_perturb(f, C.c_double(rotation_sigma), C.c_double(translation_sigma), C.c_double(point_sigma),
                                        C.c_int32(random_seed)))
            if init_depth_threshold > 0:  # BalDatasetOptions::init_depth_threshold -> filter_obs (bal_problem.cpp:471-505, :826)
                check(L.rba_bal_filter_obs(f, C.c_double(init_depth_threshold)))
            nc, nl, nobs = C.c_int32(), C.c_int32(), C.c_int64()
            check(L.rba_bal_dims(f, C.byref(nc), C.byref(nl), C.byref(nobs)))
            cams, lms = np.empty((nc.value, 10)), np.empty((nl.value, 3))
            off, oc, xy = np.empty(nl.value + 1, np.int64), np.empty(nobs.value, np.int32), np.empty((nobs.value, 2))
            check(L.rba_bal_copy(f, _p(cams), _p(lms), _p(off), _p(oc), _p(xy)))
            t = (C.c_double * 5)()
            check(L.rba_bal_load_timings(f, t))
        finally:
            L.rba_bal_free(f)
        bp = cls(cams, lms, off, oc, xy, dtype)
        bp.load_timings = dict(zip(("read", "count", "parse", "csr", "normalize"), t))
        return bp

    def num_cameras(self): return self.cams.shape[0]
    def num_landmarks(self): return self.lms.shape[0]
    def num_observations(self): return self.obs_cam.shape[0]

    # BalProblem::backup / restore (bal/bal_problem.cpp:590-608); forwarded to the device copy when attached
    def backup(self):
        if self._linearizor is not None:
            self._linearizor._backup()
        else:
            self._cams_backup[:] = self.cams
            self._lms_backup[:] = self.lms

    def restore(self):
        if self._linearizor is not None:
            self._linearizor._restore()
        else:
            self.cams[:] = self._cams_backup
            self.lms[:] = self._lms_backup

    def sync_from_device(self):
        if self._linearizor is not None:
            self._linearizor.download_state()


class LinearizorQR:
    """rootba::LinearizorQR on the GPU.  Protocol (linearizor.hpp:56-82):
    create once; per LM iteration start_iteration -> compute_error -> linearize ->
    { solve(lambda) -> [bal_problem.backup()] -> apply -> compute_error -> (restore on reject) }+ ."""

    def __init__(self, bal_problem: BalProblem, options: SolverOptions, summary: dict | None = None):
        if options.solver_type not in ("SQUARE_ROOT", "SCHUR_COMPLEMENT", "POWER_SCHUR_COMPLEMENT"):
            raise ValueError(f"solver_type {options.solver_type} is not provided by rootba_b200")
        self.bal_problem = bal_problem
        self.options = options
        self.summary = summary
        self.it_summary = None
        self.dtype = bal_problem.dtype
        self.sfx = "f32" if self.dtype == np.float32 else "f64"
        self.S = C.c_float if self.dtype == np.float32 else C.c_double
        L = _lib.lib()
        o = SolverOpts()
        L.rba_default_solver_opts(C.byref(o))
        o.use_householder_marginalization = int(options.use_householder_marginalization)
        o.use_valid_projections_only = int(options.use_projection_validity_check())
        o.robust_norm = {"NONE": 0, "HUBER": 1}[options.residual.robust_norm]
        o.huber_parameter = options.residual.huber_parameter
        o.jacobi_scaling_epsilon = options.jacobi_scaling_epsilon
        o.preconditioner_type = {"JACOBI": 0, "SCHUR_JACOBI": 1}[options.preconditioner_type]
        o.min_linear_solver_iterations = options.min_linear_solver_iterations
        o.max_linear_solver_iterations = options.max_linear_solver_iterations
        o.eta = options.eta
        o.device, o.rank, o.nranks = options.device, options.rank, options.nranks
        o.pcg_check_period = options.pcg_check_period
        o.operator_form = {"DENSE": 0, "IMPLICIT": 1}[options.operator_form]
        o.stage2_form = {"PANEL": 0, "IDENTITY": 1}[options.stage2_form]
        o.solver_type = {"SQUARE_ROOT": 0, "SCHUR_COMPLEMENT": 1, "POWER_SCHUR_COMPLEMENT": 2}[options.solver_type]  # linearizor.cpp:48-65
        o.power_order = options.power_order
        self._opts = o
        pv = ProblemView(bal_problem.num_cameras(), bal_problem.num_landmarks(), bal_problem.num_observations(),
                         bal_problem.lm_off.ctypes.data, bal_problem.obs_cam.ctypes.data, bal_problem.obs_xy.ctypes.data)
        self.h = C.c_void_p()
        check(getattr(L, f"rba_create_{self.sfx}")(C.byref(pv), C.byref(o), C.byref(self.h)))
        self.nc = bal_problem.num_cameras()
        self.nl = bal_problem.num_landmarks()
        self.upload_state()
        bal_problem._linearizor = self
        self.last_cg = CgSummary()

    # factory like Linearizor::create (linearizor.cpp:47-65)
    @staticmethod
    def create(bal_problem: BalProblem, options: SolverOptions, summary: dict | None = None) -> "LinearizorQR":
        return LinearizorQR(bal_problem, options, summary)

    def close(self):
        if getattr(self, "h", None):
            _lib.lib().rba_destroy(self.h)
            self.h = None
            if self.bal_problem._linearizor is self:
                self.bal_problem._linearizor = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- state transfer ----
    def upload_state(self):
        check(_lib.lib().rba_set_state(self.h, _p(self.bal_problem.cams), _p(self.bal_problem.lms)))

    def download_state(self):
        check(_lib.lib().rba_get_state(self.h, _p(self.bal_problem.cams), _p(self.bal_problem.lms)))

    def _backup(self):
        check(_lib.lib().rba_backup(self.h))

    def _restore(self):
        check(_lib.lib().rba_restore(self.h))

    def comm_init(self, unique_id: bytes):
        buf = C.create_string_buffer(unique_id, 128)
        check(_lib.lib().rba_comm_init(self.h, buf))

    def ipc_export(self) -> bytes:
        buf = C.create_string_buffer(128)
        check(_lib.lib().rba_ipc_export(self.h, buf))
        return buf.raw

    def ipc_import(self, all_handles: bytes):
        buf = C.create_string_buffer(all_handles, len(all_handles))
        check(_lib.lib().rba_ipc_import(self.h, buf))

    # ---- Linearizor interface ----
    def start_iteration(self, it_summary: dict | None = None):
        self.it_summary = it_summary

    def finish_iteration(self):
        pass

    def compute_error(self) -> dict:
        ri = ResidualInfo()
        check(_lib.lib().rba_compute_error(self.h, C.byref(ri)))
        if self.it_summary is not None:
            self.it_summary["residual_evaluation_time"] = self.it_summary.get("residual_evaluation_time", 0.0) + self.timings()["residual_evaluation_time"]
        return {"all": {"num_obs": ri.all_num_obs, "error": ri.all_error, "residual_sum": ri.all_residual_sum},
                "valid": {"num_obs": ri.valid_num_obs, "error": ri.valid_error, "residual_sum": ri.valid_residual_sum},
                "is_numerically_valid": bool(ri.is_numerically_valid)}

    def linearize(self):
        rc = check(_lib.lib().rba_linearize(self.h), allow_numerical_failure=True)
        if rc != 0:
            raise RbaError(rc, "did not expect numerical failure during linearization")  # linearizor_qr.cpp:121-122
        if self.it_summary is not None:
            self.it_summary["stage1_time"] = self.timings()["stage1_time"]

    def solve(self, lam: float, to_host: bool = True):
        inc = np.empty(9 * self.nc, dtype=self.dtype) if to_host else None
        cg = CgSummary()
        check(getattr(_lib.lib(), f"rba_solve_{self.sfx}")(self.h, self.S(lam), None if inc is None else _p(inc), C.byref(cg)))
        self.last_cg = cg
        if self.it_summary is not None:
            t = self.timings()
            self.it_summary.update(stage2_time=t["stage2_time"], compute_preconditioner_time=t["compute_preconditioner_time"],
                                   solve_reduced_system_time=t["solve_reduced_system_time"],
                                   linear_solver_iterations=cg.num_iterations, linear_solver_termination=cg.termination_type)
        return inc

    def apply(self, inc=None) -> float:
        l = self.S(0)
        arr = None if inc is None else np.ascontiguousarray(inc, dtype=self.dtype)
        p = None if arr is None else _p(arr)
        check(getattr(_lib.lib(), f"rba_apply_{self.sfx}")(self.h, p, C.byref(l)), allow_numerical_failure=True)
        if self.it_summary is not None:
            t = self.timings()
            self.it_summary.update(back_substitution_time=t["back_substitution_time"], update_cameras_time=t["update_cameras_time"])
        return float(l.value)

    def lm_step(self, lam: float, linearize_first: bool) -> dict:
        """one LM inner iteration with a single host synchronisation (rba_lm_step): [linearize] + solve + backup + apply +
        compute_error.  Returns the pieces optimize_lm_ours needs; the caller restores on a rejected / failed step."""
        r = LmStepResult()
        check(getattr(_lib.lib(), f"rba_lm_step_{self.sfx}")(self.h, int(linearize_first), self.S(lam), C.byref(r)),
              allow_numerical_failure=True)
        self.last_cg = r.cg
        ri = r.cost
        return {"solve_failed": bool(r.solve_failed), "l_diff": float(r.l_diff),
                "cost": {"all": {"num_obs": ri.all_num_obs, "error": ri.all_error, "residual_sum": ri.all_residual_sum},
                         "valid": {"num_obs": ri.valid_num_obs, "error": ri.valid_error, "residual_sum": ri.valid_residual_sum},
                         "is_numerically_valid": bool(ri.is_numerically_valid)}}

    def lm_run(self, max_steps: int, options: "SolverOptions | None" = None):
        """optimize_lm_ours natively (rba_lm_run): a NEW solve from the current device state, until the reference's stopping rule
        or `max_steps` iterations.  Returns (list of per-iteration dicts, terminated, phase totals in seconds)."""
        o = options or self.options
        lo = LmOpts()
        _lib.lib().rba_default_lm_opts(C.byref(lo))
        lo.initial_trust_region_radius, lo.min_trust_region_radius = o.initial_trust_region_radius, o.min_trust_region_radius
        lo.max_trust_region_radius, lo.min_relative_decrease = o.max_trust_region_radius, o.min_relative_decrease
        lo.initial_vee, lo.vee_factor, lo.function_tolerance = o.initial_vee, o.vee_factor, o.function_tolerance
        lo.max_num_iterations = o.max_num_iterations
        lo.optimized_cost = {"ERROR": 0, "ERROR_VALID": 1, "ERROR_VALID_AVG": 2}[o.optimized_cost]
        log = (LmIteration * max(max_steps, 1))()
        done, term, tot = C.c_int32(), C.c_int32(), StageTimings()
        check(getattr(_lib.lib(), f"rba_lm_run_{self.sfx}")(self.h, C.byref(lo), int(max_steps), log, C.byref(done), C.byref(term), C.byref(tot)),
              allow_numerical_failure=True)
        its = [{"lambda": log[k].lam, "cost": log[k].cost, "l_diff": log[k].l_diff, "relative_decrease": log[k].relative_decrease,
                "device_seconds": log[k].device_seconds, "cg_iterations": log[k].cg_iterations, "cg_termination": log[k].cg_termination,
                "accepted": bool(log[k].accepted), "terminated": bool(log[k].terminated)} for k in range(done.value)]
        return its, bool(term.value), struct_to_dict(tot)

    # ---- LinearizationQR-level access (tests) ----
    def timings(self) -> dict:
        t = StageTimings()
        check(_lib.lib().rba_get_timings(self.h, C.byref(t)))
        return struct_to_dict(t)

    def stats(self) -> dict:
        s = WorkloadStats()
        check(_lib.lib().rba_get_workload_stats(self.h, C.byref(s)))
        return struct_to_dict(s)

    def get_jacobian_scaling(self):
        s, d = np.empty(9 * self.nc, self.dtype), np.empty(9 * self.nc, self.dtype)
        check(_lib.lib().rba_get_jacobian_scaling(self.h, _p(s), _p(d)))
        return s, d

    def get_rhs(self):
        b = np.empty(9 * self.nc, self.dtype)
        check(_lib.lib().rba_get_rhs(self.h, _p(b)))
        return b

    def get_preconditioner(self):
        inv, blk = np.empty(81 * self.nc, self.dtype), np.empty(81 * self.nc, self.dtype)
        check(_lib.lib().rba_get_preconditioner(self.h, _p(inv), _p(blk)))
        return inv.reshape(self.nc, 9, 9), blk.reshape(self.nc, 9, 9)

    def right_multiply(self, x):
        x = np.ascontiguousarray(x, dtype=self.dtype)
        y = np.empty_like(x)
        check(_lib.lib().rba_right_multiply(self.h, _p(x), _p(y)))
        return y

    def back_substitute(self, pose_inc) -> float:
        pose_inc = np.ascontiguousarray(pose_inc, dtype=self.dtype)
        l = self.S(0)
        check(getattr(_lib.lib(), f"rba_back_substitute_{self.sfx}")(self.h, _p(pose_inc), C.byref(l)),
              allow_numerical_failure=True)
        return float(l.value)

    def debug_get_block(self, lm: int):
        n = int(self.bal_problem.lm_off[lm + 1] - self.bal_problem.lm_off[lm])
        pad = (4 - (9 * n) % 4) % 4
        rows, cols = 2 * n + 3, 9 * n + pad + 4
        out = np.zeros((rows, cols), self.dtype)
        jls = np.zeros(3, self.dtype)
        check(_lib.lib().rba_debug_get_block(self.h, C.c_int32(lm), _p(out), rows, cols, _p(jls)))
        return out, 9 * n + pad, 9 * n + pad + 3, jls

    def timer_start(self):
        check(_lib.lib().rba_timer_start(self.h))

    def timer_stop(self) -> float:
        s = C.c_double()
        check(_lib.lib().rba_timer_stop(self.h, C.byref(s)))
        return s.value

    def time_matvec(self, reps: int = 20) -> float:
        s = C.c_double()
        check(_lib.lib().rba_time_matvec(self.h, reps, C.byref(s)))
        return s.value


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    check(_lib.lib().rba_nccl_unique_id(buf))
    return buf.raw


def partition_landmarks(lm_off: np.ndarray, nranks: int) -> np.ndarray:
    lm_off = np.ascontiguousarray(lm_off, dtype=np.int64)
    bounds = np.zeros(nranks + 1, dtype=np.int32)
    check(_lib.lib().rba_partition_landmarks(C.c_int32(lm_off.shape[0] - 1), _p(lm_off), C.c_int32(nranks), _p(bounds)))
    return bounds


def _p(a):
    return C.c_void_p(a.ctypes.data)


def _cost(ri: dict, optimized_cost: str) -> float:
    if optimized_cost == "ERROR":
        return ri["all"]["error"]
    if optimized_cost == "ERROR_VALID":
        return ri["valid"]["error"]
    n = ri["valid"]["num_obs"]
    return ri["valid"]["error"] / n if n > 0 else 0.0


def bundle_adjust_manual(bal_problem: BalProblem, solver_options: SolverOptions, linearizor=None, verbose=False,
                         comm_setup=None) -> dict:
    """rootba::bundle_adjust_manual -> optimize_lm_ours (solver/bal_bundle_adjustment.cpp:249-544):
    the host-serial LM trust-region loop, unchanged, driving a Linearizor."""
    o = solver_options
    S = np.float32 if bal_problem.dtype == np.float32 else np.float64
    min_lambda = S(1.0 / o.max_trust_region_radius)
    max_lambda = S(1.0 / o.min_trust_region_radius)
    vee_factor, initial_vee = S(o.vee_factor), S(o.initial_vee)
    lam = S(1.0 / o.initial_trust_region_radius)
    lambda_vee = initial_vee
    summary = {"iterations": [], "num_linear_solves": 0, "num_residual_evaluations": 0, "num_jacobian_evaluations": 0,
               "termination_type": "NO_CONVERGENCE", "message": ""}
    t_total = time.perf_counter()
    own = linearizor is None
    if own:
        linearizor = LinearizorQR.create(bal_problem, o, summary)
        if comm_setup is not None:
            comm_setup(linearizor)
    summary["preprocessor_time"] = time.perf_counter() - t_total
    t_iter = [time.perf_counter()]

    def log_iteration(s):  # finish_iteration (bal_bundle_adjustment.cpp:56-88): wall-clock stamps, then push
        now = time.perf_counter()
        s["iteration_time"], s["cumulative_time"] = now - t_iter[0], now - t_total
        t_iter[0] = now
        summary["iterations"].append(s)
    terminated = False
    it = 0
    max_lm_iter = o.max_num_iterations
    try:
        _lm_loop(bal_problem, o, S, linearizor, summary, log_iteration, min_lambda, max_lambda, vee_factor, initial_vee, lam, lambda_vee,
                 max_lm_iter, verbose)
    except BaseException:
        if own:
            linearizor.close()  # do not leak the device handle when the loop raises (numerical failure during linearisation ...)
        raise
    bal_problem.sync_from_device()
    summary["total_time"] = time.perf_counter() - t_total
    summary["minimizer_time"] = summary["total_time"] - summary["preprocessor_time"]
    if own:
        summary["stats"] = linearizor.stats()
        linearizor.close()
    return summary


def _lm_loop(bal_problem, o, S, linearizor, summary, log_iteration, min_lambda, max_lambda, vee_factor, initial_vee, lam, lambda_vee,
             max_lm_iter, verbose):
    """the body of optimize_lm_ours (solver/bal_bundle_adjustment.cpp:291-521)"""
    terminated = False
    it = 0
    with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
        while it <= max_lm_iter and not terminated:
            it_summary = {"iteration": it}
            linearizor.start_iteration(it_summary)
            ri = linearizor.compute_error()
            summary["num_residual_evaluations"] += 1
            if not ri["is_numerically_valid"]:
                raise RbaError(1, "did not expect numerical failure during linearization")  # :307-308
            if it == 0:
                linearizor.finish_iteration()
                it_summary.update(cost=ri, trust_region_radius=1 / float(lam), step_is_successful=True, step_is_valid=True,
                                  lam=float(lam))
                log_iteration(it_summary)
                it += 1
                continue
            linearizor.linearize()
            summary["num_jacobian_evaluations"] += 1
            j = 0
            while it <= max_lm_iter and not terminated:
                if j > 0:
                    it_summary = {"iteration": it}
                    linearizor.start_iteration(it_summary)
                j += 1
                inc = linearizor.solve(float(lam))
                summary["num_linear_solves"] += 1
                it_summary["lam"] = float(lam)
                if not np.all(np.isfinite(inc)):
                    it_summary.update(step_is_valid=False, step_is_successful=False)
                    lam = S(lambda_vee * lam)
                    lambda_vee = S(lambda_vee * vee_factor)
                    linearizor.finish_iteration()
                    it_summary["trust_region_radius"] = 1 / float(lam)
                    log_iteration(it_summary)
                    it += 1
                    if lam > max_lambda:
                        terminated = True
                        summary["message"] = "Solver did not converge and reached maximum damping lambda"
                    continue
                bal_problem.backup()
                l_diff = S(linearizor.apply(inc))
                ri2 = linearizor.compute_error()
                summary["num_residual_evaluations"] += 1
                it_summary["cost"] = ri2
                it_summary["l_diff"] = float(l_diff)
                if not math.isfinite(float(l_diff)) or not ri2["is_numerically_valid"]:
                    it_summary.update(step_is_valid=False, step_is_successful=False)
                else:
                    f_diff = S(_cost(ri, o.optimized_cost) - _cost(ri2, o.optimized_cost))
                    if o.optimized_cost == "ERROR_VALID_AVG":
                        l_diff = S(l_diff / ri["valid"]["num_obs"])
                    step_quality = S(f_diff / l_diff)
                    it_summary["relative_decrease"] = float(step_quality)
                    it_summary["step_is_valid"] = bool(l_diff > 0)
                    it_summary["step_is_successful"] = bool(it_summary["step_is_valid"] and step_quality > o.min_relative_decrease)
                if it_summary["step_is_successful"]:
                    lam = S(lam * S(max(1.0 / 3, 1 - (2 * it_summary["relative_decrease"] - 1) ** 3)))
                    lam = max(min_lambda, lam)
                    lambda_vee = initial_vee
                    linearizor.finish_iteration()
                    it_summary["trust_region_radius"] = 1 / float(lam)
                    zero = {"all": {"num_obs": 0, "error": 0.0}, "valid": {"num_obs": 0, "error": 0.0}}
                    prev = _cost(summary["iterations"][-1].get("cost", zero), "ERROR" if o.optimized_cost == "ERROR" else "ERROR_VALID")
                    cur = _cost(ri2, "ERROR" if o.optimized_cost == "ERROR" else "ERROR_VALID")
                    log_iteration(it_summary)
                    it += 1
                    if abs(prev - cur) <= o.function_tolerance * cur:  # :174-201
                        terminated = True
                        summary["termination_type"] = "CONVERGENCE"
                        summary["message"] = "Function tolerance reached."
                    if verbose:
                        print(f"  it {it - 1}: cost {cur:.6e} lambda {float(lam):.1e} cg {it_summary.get('linear_solver_iterations')}")
                    break
                else:
                    lam = S(lambda_vee * lam)
                    lambda_vee = S(lambda_vee * vee_factor)
                    linearizor.finish_iteration()
                    it_summary["trust_region_radius"] = 1 / float(lam)
                    log_iteration(it_summary)
                    bal_problem.restore()
                    it += 1
                    if lam > max_lambda:
                        terminated = True
                        summary["message"] = "Solver did not converge and reached maximum damping lambda"
    if not terminated:
        summary["message"] = f"Solver did not converge after maximum number of {max_lm_iter} iterations"

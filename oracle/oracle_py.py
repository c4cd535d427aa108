"""ctypes binding of the CPU oracle (oracle/_build/librootba_oracle.so).

TEST INFRASTRUCTURE ONLY ("parity unpinned", see rootba_oracle.hpp).  Importable from
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs -- never
from the rootba_b200 package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "librootba_oracle.so")


def build(force: bool = False) -> str:
    if force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB_PATH)
        for f in ("rootba_oracle.hpp", "rootba_oracle_capi.cpp")
    ):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


class COptions(C.Structure):
    _fields_ = [
        ("use_householder", C.c_int), ("use_valid_projections_only", C.c_int), ("robust_norm", C.c_int),
        ("huber_parameter", C.c_double), ("jacobi_scaling_epsilon", C.c_double),
        ("preconditioner_type", C.c_int), ("min_linear_solver_iterations", C.c_int),
        ("max_linear_solver_iterations", C.c_int), ("eta", C.c_double), ("staged_execution", C.c_int),
        ("reduction_alg", C.c_int), ("max_num_iterations", C.c_int),
        ("initial_trust_region_radius", C.c_double), ("min_trust_region_radius", C.c_double),
        ("max_trust_region_radius", C.c_double), ("min_relative_decrease", C.c_double),
        ("function_tolerance", C.c_double), ("initial_vee", C.c_double), ("vee_factor", C.c_double),
        ("num_threads", C.c_int), ("optimized_cost", C.c_int), ("verbose", C.c_int),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_max_threads.restype = C.c_int
        for sfx in ("f32", "f64"):
            getattr(_lib, f"orc_create_{sfx}").restype = C.c_void_p
    return _lib


def default_options(**kw) -> COptions:
    o = COptions()
    lib().orc_default_options(C.byref(o))
    for k, v in kw.items():
        if not hasattr(o, k):
            raise AttributeError(k)
        setattr(o, k, v)
    return o


def max_threads() -> int:
    return int(lib().orc_max_threads())


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


LOG_COLUMNS = ["iteration", "cost", "cost_valid", "num_obs_valid", "step_is_valid", "step_is_successful",
               "lambda", "trust_region_radius", "relative_decrease", "l_diff", "cg_iterations",
               "stage1_time", "stage2_time", "precond_time", "pcg_time", "backsub_time", "update_time",
               "error_time", "iteration_time", "_"]


class Oracle:
    """CPU restatement of BalProblem + LinearizationQR + LinearizorQR for one problem."""

    def __init__(self, prob, dtype=np.float64, options: COptions | None = None):
        self.dtype = np.dtype(dtype)
        self.sfx = "f32" if self.dtype == np.float32 else "f64"
        self.S = C.c_float if self.dtype == np.float32 else C.c_double
        self.nc, self.nl, self.nobs = prob.nc, prob.nl, prob.nobs
        self.lm_off = np.ascontiguousarray(prob.lm_off, dtype=np.int64)
        self.obs_cam = np.ascontiguousarray(prob.obs_cam, dtype=np.int32)
        self.options = options or default_options()
        cams = np.ascontiguousarray(prob.cams, dtype=self.dtype)
        lms = np.ascontiguousarray(prob.lms, dtype=self.dtype)
        xy = np.ascontiguousarray(prob.obs_xy, dtype=self.dtype)
        self.h = C.c_void_p(self._f("create")(C.c_int(self.nc), C.c_int(self.nl), _p(self.lm_off),
                                              _p(self.obs_cam), _p(xy), _p(cams), _p(lms),
                                              C.byref(self.options)))

    def _f(self, name):
        return getattr(lib(), f"orc_{name}_{self.sfx}")

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self._f("destroy")(self.h)
                self.h = None
        except Exception:
            pass

    def _vec(self, n):
        return np.zeros(n, dtype=self.dtype)

    def set_options(self, options: COptions):
        self.options = options
        self._f("set_options")(self.h, C.byref(options))

    # --- state ---
    def get_state(self):
        cams, lms = self._vec(10 * self.nc), self._vec(3 * self.nl)
        self._f("get_state")(self.h, _p(cams), _p(lms))
        return cams.reshape(self.nc, 10), lms.reshape(self.nl, 3)

    def set_state(self, cams, lms):
        cams = np.ascontiguousarray(cams, dtype=self.dtype)
        lms = np.ascontiguousarray(lms, dtype=self.dtype)
        self._f("set_state")(self.h, _p(cams), _p(lms))

    def backup(self):
        self._f("backup")(self.h)

    def restore(self):
        self._f("restore")(self.h)

    def compute_error(self) -> dict:
        out = np.zeros(7)
        self._f("compute_error")(self.h, _p(out))
        return {"all": {"num_obs": int(out[0]), "error": out[1], "residual_sum": out[2]},
                "valid": {"num_obs": int(out[3]), "error": out[4], "residual_sum": out[5]},
                "is_numerically_valid": bool(out[6])}

    # --- LinearizationQR level ---
    def stage1(self, jacobi_blocks=False):
        d = self._vec(9 * self.nc)
        rc = self._f("stage1")(self.h, _p(d), C.c_int(int(jacobi_blocks)))
        return d, rc == 0

    def set_pose_damping(self, lam):
        self._f("set_pose_damping")(self.h, self.S(lam))

    def stage2(self, lam, scaling=None, schur_blocks=True):
        b = self._vec(9 * self.nc)
        blocks = self._vec(81 * self.nc)
        sc = None if scaling is None else _p(np.ascontiguousarray(scaling, dtype=self.dtype))
        self._f("stage2")(self.h, self.S(lam), sc, C.c_int(int(schur_blocks)), _p(b), _p(blocks))
        return b, blocks.reshape(self.nc, 9, 9)

    def right_multiply(self, x):
        x = np.ascontiguousarray(x, dtype=self.dtype)
        y = self._vec(9 * self.nc)
        self._f("right_multiply")(self.h, _p(x), _p(y))
        return y

    def back_substitute(self, pose_inc):
        pose_inc = np.ascontiguousarray(pose_inc, dtype=self.dtype)
        l = self.S(0)
        rc = self._f("back_substitute")(self.h, _p(pose_inc), C.byref(l))
        return float(l.value), rc == 0

    def get_block(self, lm: int):
        r, c, li, ri = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        self._f("block_dims")(self.h, C.c_int(lm), C.byref(r), C.byref(c), C.byref(li), C.byref(ri))
        out = self._vec(r.value * c.value)
        sc = self._vec(3)
        self._f("get_block")(self.h, C.c_int(lm), _p(out), _p(sc))
        return out.reshape(r.value, c.value), li.value, ri.value, sc

    # --- LinearizorQR level ---
    def linearize(self) -> bool:
        return self._f("linearize")(self.h) == 0

    def get_scaling(self):
        s = self._vec(9 * self.nc)
        self._f("get_scaling")(self.h, _p(s))
        return s

    def solve(self, lam, want_debug=False):
        inc = self._vec(9 * self.nc)
        b = self._vec(9 * self.nc)
        inv = self._vec(81 * self.nc)
        it, term = C.c_int(), C.c_int()
        self._f("solve")(self.h, self.S(lam), _p(inc), _p(b), _p(inv), C.byref(it), C.byref(term))
        if want_debug:
            return inc, {"b": b, "inv_blocks": inv.reshape(self.nc, 9, 9), "cg_iterations": it.value,
                         "cg_termination": term.value}
        return inc, it.value

    def apply(self, inc) -> float:
        inc = np.ascontiguousarray(inc, dtype=self.dtype)
        l = self.S(0)
        self._f("apply")(self.h, _p(inc), C.byref(l))
        return float(l.value)

    def timings(self) -> dict:
        t = np.zeros(8)
        self._f("get_timings")(self.h, _p(t))
        return dict(zip(["stage1", "stage2", "precond", "pcg", "backsub", "update", "error", "matvecs"], t))

    def optimize(self, max_rows=256):
        log = np.zeros((max_rows, 20))
        term = C.c_int()
        n = self._f("optimize")(self.h, _p(log), C.c_int(max_rows), C.byref(term))
        rows = [dict(zip(LOG_COLUMNS, log[i])) for i in range(n)]
        return rows, term.value

    # --- LinearizorSC / LinearizorPowerSC (solver/linearizor_sc.cpp, solver/linearizor_power_sc.cpp) ---
    def scl_linearize(self):
        self._f("scl_linearize")(self.h)

    def scl_get_scaling(self):
        s = self._vec(9 * self.nc)
        self._f("scl_get_scaling")(self.h, _p(s))
        return s

    def scl_solve(self, lam):
        inc, b, inv = self._vec(9 * self.nc), self._vec(9 * self.nc), self._vec(81 * self.nc)
        it, term = C.c_int(), C.c_int()
        self._f("scl_solve")(self.h, self.S(lam), _p(inc), _p(b), _p(inv), C.byref(it), C.byref(term))
        return inc, {"b": b, "inv_blocks": inv.reshape(self.nc, 9, 9), "cg_iterations": it.value, "cg_termination": term.value}

    def scl_power_solve(self, lam, power_order=20, q_tolerance=0.1):
        inc, b = self._vec(9 * self.nc), self._vec(9 * self.nc)
        it, term = C.c_int(), C.c_int()
        self._f("scl_power_solve")(self.h, self.S(lam), C.c_int(power_order), self.S(q_tolerance), _p(inc), _p(b), C.byref(it), C.byref(term))
        return inc, {"b": b, "power_order": it.value, "termination": term.value}

    def scl_e0(self, lam, x):
        x = np.ascontiguousarray(x, dtype=self.dtype)
        y = self._vec(9 * self.nc)
        self._f("scl_e0")(self.h, self.S(lam), _p(x), _p(y))
        return y

    def scl_apply(self, inc) -> float:
        inc = np.ascontiguousarray(inc, dtype=self.dtype)
        l = self.S(0)
        self._f("scl_apply")(self.h, _p(inc), C.byref(l))
        return float(l.value)

    # --- Schur complement cross-check ---
    def sc_linearize(self):
        d = self._vec(9 * self.nc)
        self._f("sc_linearize")(self.h, _p(d))
        return d

    def sc_scale_Jp(self, scaling):
        scaling = np.ascontiguousarray(scaling, dtype=self.dtype)
        self._f("sc_scale_Jp")(self.h, _p(scaling))

    def sc_get_Hb(self, lam, pose_damping, x):
        x = np.ascontiguousarray(x, dtype=self.dtype)
        b, d, y = self._vec(9 * self.nc), self._vec(81 * self.nc), self._vec(9 * self.nc)
        self._f("sc_get_Hb")(self.h, self.S(lam), self.S(pose_damping), _p(x), _p(b), _p(d), _p(y))
        return b, d.reshape(self.nc, 9, 9), y

    def sc_back_substitute(self, lam, pose_inc):
        pose_inc = np.ascontiguousarray(pose_inc, dtype=self.dtype)
        l = self.S(0)
        lms = self._vec(3 * self.nl)
        self._f("sc_back_substitute")(self.h, self.S(lam), _p(pose_inc), C.byref(l), _p(lms))
        return float(l.value), lms.reshape(self.nl, 3)


# --- free functions ---
def linearize_point(obs, p_w, cam, dtype=np.float64):
    dt = np.dtype(dtype)
    sfx = "f32" if dt == np.float32 else "f64"
    obs, p_w, cam = (np.ascontiguousarray(a, dtype=dt) for a in (obs, p_w, cam))
    res, Jp, Ji, Jl = np.zeros(2, dt), np.zeros(12, dt), np.zeros(6, dt), np.zeros(6, dt)
    valid = getattr(lib(), f"orc_linearize_point_{sfx}")(_p(obs), _p(p_w), _p(cam), _p(res), _p(Jp), _p(Ji), _p(Jl))
    return res, Jp.reshape(2, 6), Ji.reshape(2, 3), Jl.reshape(2, 3), bool(valid)


def camera_apply_inc(cam, inc, dtype=np.float64):
    dt = np.dtype(dtype)
    sfx = "f32" if dt == np.float32 else "f64"
    cam = np.array(cam, dtype=dt)
    inc = np.ascontiguousarray(inc, dtype=dt)
    getattr(lib(), f"orc_camera_apply_inc_{sfx}")(_p(cam), _p(inc))
    return cam


def invert_block9(block, diag=None, dtype=np.float64):
    dt = np.dtype(dtype)
    sfx = "f32" if dt == np.float32 else "f64"
    block = np.ascontiguousarray(block, dtype=dt)
    out = np.zeros(81, dt)
    d = None if diag is None else _p(np.ascontiguousarray(diag, dtype=dt))
    getattr(lib(), f"orc_invert_block9_{sfx}")(_p(block), d, _p(out))
    return out.reshape(9, 9)


def load_bal(path: str, normalize=True, scale=100.0, init_depth_threshold=0.0):
    """BAL text loader + default normalisation (double) + filter_obs (bal_problem.cpp:471-505).  Returns a dict of arrays."""
    nc, nl, nobs = C.c_int(), C.c_int(), C.c_int64()
    rc = lib().orc_bal_load(path.encode(), C.c_int(int(normalize)), C.c_double(scale), C.byref(nc),
                            C.byref(nl), C.byref(nobs))
    if rc != 0:
        raise RuntimeError(f"orc_bal_load failed rc={rc}")
    if init_depth_threshold > 0:
        lib().orc_bal_filter_obs(C.c_double(init_depth_threshold), C.byref(nl), C.byref(nobs))
    cams, lms = np.zeros((nc.value, 10)), np.zeros((nl.value, 3))
    off, oc, xy = np.zeros(nl.value + 1, np.int64), np.zeros(nobs.value, np.int32), np.zeros((nobs.value, 2))
    lib().orc_bal_get(_p(cams), _p(lms), _p(off), _p(oc), _p(xy))
    return {"cams": cams, "lms": lms, "lm_off": off, "obs_cam": oc, "obs_xy": xy}


def normalize(cams, lms, scale=100.0):
    cams = np.array(cams, dtype=np.float64)
    lms = np.array(lms, dtype=np.float64)
    lib().orc_normalize(C.c_int(cams.shape[0]), C.c_int(lms.shape[0]), _p(cams), _p(lms), C.c_double(scale))
    return cams, lms

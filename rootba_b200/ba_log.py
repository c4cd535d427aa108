"""ba_log.json in the reference's format for runs of the Python host (same layout as rootba_b200/host/ba_log.hpp writes for
`bal_qr`): one flat JSON object, every member of BaLog::BaIteration (src/rootba/bal/ba_log.hpp:139-237) a top-level array
with one entry per logged iteration, plus "_type": "rootba" and "_static": {problem_info, timing, solver}
(src/rootba/bal/ba_log.cpp:62-150).  Values follow log_summary (src/rootba/bal/ba_log_utils.cpp:40-166) and
finish_iteration / finish_solve (src/rootba/solver/bal_bundle_adjustment.cpp:56-140)."""
from __future__ import annotations

import json
import math

import numpy as np

# BaLog::BaIteration members in declaration order
BA_ITERATION_FIELDS = (
    "iteration", "linear_solver_type", "step_is_valid", "step_is_nonmonotonic", "step_is_successful", "num_obs", "num_obs_valid",
    "num_obs_valid_change", "cost", "cost_change", "cost_valid", "cost_valid_change", "cost_avg_valid", "cost_avg_valid_change",
    "grad_projected_norm", "grad_projected_max_norm", "grad_norm", "grad_max_norm", "residual_block_mean",
    "residual_block_valid_mean", "step_norm", "relative_decrease", "trust_region_radius", "linear_solver_iterations",
    "iteration_time", "cumulative_time", "logging_time", "step_solver_time", "residual_evaluation_time", "jacobian_evaluation_time",
    "scale_landmark_jacobian_time", "perform_qr_time", "stage1_time", "scale_pose_jacobian_time", "landmark_damping_time",
    "compute_preconditioner_time", "compute_gradient_time", "stage2_time", "prepare_time", "solve_reduced_system_time",
    "back_substitution_time", "update_cameras_time", "resident_memory", "resident_memory_peak")
_INT_FIELDS = {"iteration", "num_obs", "num_obs_valid", "num_obs_valid_change", "linear_solver_iterations", "resident_memory",
               "resident_memory_peak"}
_BOOL_FIELDS = {"step_is_valid", "step_is_nonmonotonic", "step_is_successful"}
_TERMINATION = {"CONVERGENCE": 0, "NO_CONVERGENCE": 1, "FAILURE": 2}  # enum TerminationType, solver/solver_summary.hpp:83-98


def _num(v):
    v = float(v)
    return v if math.isfinite(v) else None  # nlohmann::json writes NaN / inf as null


def _mean(item):
    return item["residual_sum"] / item["num_obs"] if item.get("num_obs", 0) > 0 else 0.0


def _avg(item):
    return item["error"] / item["num_obs"] if item.get("num_obs", 0) > 0 else 0.0


def summarize_problem(bal_problem, input_path: str = "", compute_sparsity: bool = True) -> dict:
    """BalProblem::summarize_problem + compute_rcs_sparsity (bal/bal_problem.cpp:648-759)"""
    n = np.diff(bal_problem.lm_off).astype(np.float64)
    nc = int(bal_problem.num_cameras())
    stats = {"mean": float(n.mean()), "min": float(n.min()), "max": float(n.max()), "stddev": float(np.sqrt(((n - n.mean()) ** 2).mean()))}
    sparsity = 0.0
    if compute_sparsity and 0 < nc and nc * nc < 2 ** 31:
        mask = np.zeros((nc, nc), dtype=bool)
        off, cam = bal_problem.lm_off, bal_problem.obs_cam
        for l in range(len(off) - 1):
            c = cam[off[l]:off[l + 1]]
            mask[np.ix_(c, c)] = True
        lower = int(np.tril(mask, -1).sum())
        sparsity = 1.0 - (nc + 2 * lower) / float(nc * nc)
    return {"type": "bal", "input_path": input_path, "num_cameras": nc, "num_landmarks": int(bal_problem.num_landmarks()),
            "num_observations": int(bal_problem.num_observations()), "rcs_sparsity": sparsity, "per_lm_obs": stats,
            "per_host_lms": {"mean": 0.0, "min": 0.0, "max": 0.0, "stddev": 0.0}}


def make_ba_log(summary: dict, problem_info: dict, timing: dict | None = None, solver_type: str = "bal_qr") -> dict:
    """`summary` as returned by rootba_b200.bundle_adjust_manual"""
    cols = {k: [] for k in BA_ITERATION_FIELDS}
    prev_row, prev_cost = None, None
    lin_t = res_t = 0.0
    n_ok, n_bad = -1, 0  # finish_solve: iteration 0 is not a step
    for it in summary["iterations"]:
        row = {k: (False if k in _BOOL_FIELDS else 0 if k in _INT_FIELDS else "" if k == "linear_solver_type" else 0.0) for k in BA_ITERATION_FIELDS}
        ok = bool(it.get("step_is_successful", False))
        row.update(iteration=int(it["iteration"]), linear_solver_type="bal_qr" if it["iteration"] > 0 else "",
                   step_is_valid=bool(it.get("step_is_valid", False)), step_is_successful=ok)
        cost = it.get("cost")
        if (ok or prev_row is None) and cost is not None:
            a, v = cost["all"], cost["valid"]
            row.update(num_obs=int(a["num_obs"]), num_obs_valid=int(v["num_obs"]), cost=_num(a["error"]), cost_valid=_num(v["error"]),
                       cost_avg_valid=_num(_avg(v)), residual_block_mean=_num(_mean(a)), residual_block_valid_mean=_num(_mean(v)),
                       relative_decrease=_num(it.get("relative_decrease", 0.0)))
            if it["iteration"] > 0 and prev_cost is not None:  # cost.compared_to(previous logged cost), residual_info.cpp:43-53
                pa, pv = prev_cost["all"], prev_cost["valid"]
                row.update(num_obs_valid_change=int(pv["num_obs"] - v["num_obs"]), cost_change=_num(pa["error"] - a["error"]),
                           cost_valid_change=_num(pv["error"] - v["error"]), cost_avg_valid_change=_num(_avg(pv) - _avg(v)))
        elif prev_row is not None:  # unsuccessful: repeat the previous columns "for monotonic plots" (ba_log_utils.cpp:119-137)
            for k in ("num_obs", "num_obs_valid", "cost", "cost_valid", "cost_avg_valid", "residual_block_mean", "residual_block_valid_mean"):
                row[k] = prev_row[k]
        t = {k: float(it.get(k, 0.0)) for k in ("stage1_time", "stage2_time", "compute_preconditioner_time", "solve_reduced_system_time",
                                                "back_substitution_time", "update_cameras_time", "residual_evaluation_time")}
        step_solver = t["stage2_time"] + t["solve_reduced_system_time"] + t["back_substitution_time"]  # bal_bundle_adjustment.cpp:56-66
        row.update(trust_region_radius=_num(it.get("trust_region_radius", 0.0)), linear_solver_iterations=int(it.get("linear_solver_iterations", 0)),
                   iteration_time=_num(it.get("iteration_time", 0.0)), cumulative_time=_num(it.get("cumulative_time", 0.0)),
                   step_solver_time=step_solver, **t)
        lin_t += step_solver
        res_t += t["residual_evaluation_time"]
        n_ok, n_bad = (n_ok + 1, n_bad) if ok else (n_ok, n_bad + 1)
        for k in BA_ITERATION_FIELDS:
            cols[k].append(row[k])
        prev_row = row
        if cost is not None:
            prev_cost = cost
    timing = dict(timing or {})
    tm = {"load": float(timing.get("load", 0.0)), "preprocess": float(timing.get("preprocess", 0.0)),
          "optimize": float(timing.get("optimize", summary.get("total_time", 0.0))), "postprocess": float(timing.get("postprocess", 0.0))}
    tm = {"total": tm["load"] + tm["preprocess"] + tm["optimize"], **tm}
    solver = {"solver_type": solver_type, "termination_type": _TERMINATION.get(summary.get("termination_type", "FAILURE"), 2),
              "message": summary.get("message", ""), "num_successful_steps": n_ok, "num_unsuccessful_steps": n_bad,
              "logging_time_in_seconds": 0.0, "preprocessor_time_in_seconds": float(summary.get("preprocessor_time", 0.0)),
              "minimizer_time_in_seconds": float(summary.get("minimizer_time", 0.0)), "postprocessor_time_in_seconds": 0.0,
              "total_time_in_seconds": float(summary.get("total_time", 0.0)), "linear_solver_time_in_seconds": lin_t,
              "num_linear_solves": int(summary.get("num_linear_solves", 0)), "residual_evaluation_time_in_seconds": res_t,
              "num_residual_evaluations": int(summary.get("num_residual_evaluations", 0)), "jacobian_evaluation_time_in_seconds": 0.0,
              "num_jacobian_evaluations": int(summary.get("num_jacobian_evaluations", 0)), "num_threads_given": 0, "num_threads_used": 0,
              "num_threads_available": 0, "resident_memory_peak": 0}
    out = dict(cols)
    out["_type"] = "rootba"
    out["_static"] = {"problem_info": problem_info, "timing": tm, "solver": solver}
    return out


def save_ba_log(path: str, summary: dict, bal_problem, input_path: str = "", timing: dict | None = None) -> None:
    with open(path, "w") as f:
        json.dump(make_ba_log(summary, summarize_problem(bal_problem, input_path), timing), f, indent=4)

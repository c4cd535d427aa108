// ba_log.json in the reference's format, so that a run of the GPU solver can be read by the
// reference's own evaluation tooling (python/rootba/log.py: BaLog, load_ba_log; tables / plots in python/rootba).
//
// Format (src/rootba/bal/ba_log.cpp:62-150): ONE flat JSON object; every member of BaLog::BaIteration
// (src/rootba/bal/ba_log.hpp:139-237) is a top-level key holding an array with one entry per logged iteration
// ("column" layout, "for easier import in matlab"), plus "_type": "rootba" and "_static": {problem_info, timing,
// solver} (ba_log.hpp:54-136, 240-252).  Values are filled as log_summary does (src/rootba/bal/ba_log_utils.cpp:40-166):
// an unsuccessful iteration repeats the cost columns of the previous entry "for monotonic plots" (:119-137).
// Quantities the GPU path does not separate (jacobian_evaluation_time, perform_qr_time, ... are all inside stage1_time
// in staged execution, exactly as in the reference: bal_bundle_adjustment.cpp:56-61) are written as 0.
#pragma once

#include <cmath>
#include <cstdint>
#include <fstream>
#include <iomanip>
#include <sstream>
#include <string>
#include <vector>

#include "solver.hpp"

namespace rootba_b200 {

struct DatasetStats { double mean = 0, min = 0, max = 0, stddev = 0; };
struct DatasetSummary {  // bal/bal_pipeline_summary.hpp DatasetSummary -> BaLog::ProblemInfo
  std::string type = "bal", input_path;
  int num_cameras = 0, num_landmarks = 0;
  long long num_observations = 0;
  double rcs_sparsity = 0;
  DatasetStats per_lm_obs, per_host_lms;
};
struct PipelineTimingSummary { double load_time = 0, preprocess_time = 0, optimize_time = 0, postprocess_time = 0; };

// ref: BalProblem::summarize_problem + compute_rcs_sparsity (bal/bal_problem.cpp:648-759)
template <class Scalar, class Problem>
DatasetSummary summarize_problem(const Problem& p, const std::string& input_path, bool compute_sparsity = true) {
  DatasetSummary s;
  s.input_path = input_path;
  s.num_cameras = p.num_cameras();
  s.num_landmarks = p.num_landmarks();
  s.num_observations = p.num_observations();
  std::vector<int64_t> off;
  std::vector<int32_t> oc;
  std::vector<Scalar> xy;
  p.export_topology(off, oc, xy);
  const int nl = s.num_landmarks;
  double sum = 0, mn = 1e300, mx = -1e300;
  for (int l = 0; l < nl; ++l) { const double n = double(off[l + 1] - off[l]); sum += n; mn = std::min(mn, n); mx = std::max(mx, n); }
  s.per_lm_obs.mean = nl ? sum / nl : 0;
  s.per_lm_obs.min = nl ? mn : 0;
  s.per_lm_obs.max = nl ? mx : 0;
  double sq = 0;
  for (int l = 0; l < nl; ++l) { const double d = double(off[l + 1] - off[l]) - s.per_lm_obs.mean; sq += d * d; }
  s.per_lm_obs.stddev = nl ? std::sqrt(sq / nl) : 0;
  if (compute_sparsity && s.num_cameras > 0 && (long long)s.num_cameras * s.num_cameras < (1LL << 31)) {
    const long long nc = s.num_cameras;
    std::vector<char> mask((size_t)(nc * nc), 0);
    for (int l = 0; l < nl; ++l)
      for (int64_t a = off[l]; a < off[l + 1]; ++a)
        for (int64_t b = off[l]; b < a; ++b) mask[(size_t)(oc[a] * nc + oc[b])] = 1;  // cam_j < cam_i (ascending order)
    long long cnt = 0;
    for (char m : mask) cnt += m;
    s.rcs_sparsity = 1.0 - double(nc + 2 * cnt) / double(nc * nc);
  }
  return s;
}

namespace detail {
inline std::string json_str(const std::string& v) {
  std::ostringstream o;
  o << '"';
  for (char c : v) {
    if (c == '"' || c == '\\') o << '\\' << c;
    else if (c == '\n') o << "\\n";
    else if ((unsigned char)c < 0x20) o << ' ';
    else o << c;
  }
  o << '"';
  return o.str();
}
inline std::string json_num(double v) {
  if (!std::isfinite(v)) return "null";  // nlohmann::json dumps NaN / inf as null
  std::ostringstream o;
  o << std::setprecision(17) << v;
  std::string t = o.str();
  if (t.find_first_of(".eE") == std::string::npos) t += ".0";  // a double stays a JSON float (nlohmann::json prints 100.0)
  return t;
}
template <class It, class F>
void column(std::ostream& f, const char* name, const std::vector<It>& rows, F&& value, bool last = false) {
  f << "    " << json_str(name) << ": [";
  for (size_t i = 0; i < rows.size(); ++i) f << (i ? ", " : "") << value(rows[i]);
  f << "]" << (last ? "\n" : ",\n");
}
}  // namespace detail

inline int termination_type_code(const std::string& t) {  // enum TerminationType, solver/solver_summary.hpp:83-98
  if (t == "CONVERGENCE") return 0;
  if (t == "NO_CONVERGENCE") return 1;
  return 2;  // FAILURE
}

// One row per logged iteration, already in BaLog::BaIteration form
struct BaIterationRow {
  int iteration = -1;
  std::string linear_solver_type;
  bool step_is_valid = false, step_is_nonmonotonic = false, step_is_successful = false;
  long long num_obs = 0, num_obs_valid = 0, num_obs_valid_change = 0;
  double cost = 0, cost_change = 0, cost_valid = 0, cost_valid_change = 0, cost_avg_valid = 0, cost_avg_valid_change = 0;
  double residual_block_mean = 0, residual_block_valid_mean = 0, relative_decrease = 0, trust_region_radius = 0;
  int linear_solver_iterations = 0;
  double iteration_time = 0, cumulative_time = 0, step_solver_time = 0;
  double residual_evaluation_time = 0, stage1_time = 0, compute_preconditioner_time = 0, stage2_time = 0,
         solve_reduced_system_time = 0, back_substitution_time = 0, update_cameras_time = 0;
};

// log_summary(BaIteration&, prev, IterationSummary) (ba_log_utils.cpp:97-166) + finish_iteration's derived values
// (bal_bundle_adjustment.cpp:56-73: step_solver_time, cost_change = previous - this)
inline std::vector<BaIterationRow> make_ba_iterations(const SolverSummary& summary) {
  std::vector<BaIterationRow> rows;
  for (size_t i = 0; i < summary.iterations.size(); ++i) {
    const IterationSummary& it = summary.iterations[i];
    BaIterationRow r;
    r.iteration = it.iteration;
    r.linear_solver_type = it.iteration > 0 ? "bal_qr" : "";  // set by LinearizorQR::solve only (linearizor_qr.cpp:256)
    r.step_is_valid = it.step_is_valid;
    r.step_is_successful = it.step_is_successful;
    const auto mean = [](const ResidualItem& x) { return x.num_obs > 0 ? x.residual_sum / x.num_obs : 0.0; };
    if (it.step_is_successful || rows.empty()) {
      r.num_obs = it.cost.all.num_obs;
      r.num_obs_valid = it.cost.valid.num_obs;
      r.cost = it.cost.all.error;
      r.cost_valid = it.cost.valid.error;
      r.cost_avg_valid = it.cost.valid.error_avg();
      r.residual_block_mean = mean(it.cost.all);
      r.residual_block_valid_mean = mean(it.cost.valid);
      r.relative_decrease = it.relative_decrease;
      if (it.iteration > 0 && i > 0) {  // compared_to(previous iteration's cost), residual_info.cpp:43-53
        const ResidualInfo& p = summary.iterations[i - 1].cost;
        r.num_obs_valid_change = p.valid.num_obs - it.cost.valid.num_obs;
        r.cost_change = p.all.error - it.cost.all.error;
        r.cost_valid_change = p.valid.error - it.cost.valid.error;
        r.cost_avg_valid_change = p.valid.error_avg() - it.cost.valid.error_avg();
      }
    } else {
      const BaIterationRow& p = rows.back();
      r.num_obs = p.num_obs; r.num_obs_valid = p.num_obs_valid;
      r.cost = p.cost; r.cost_valid = p.cost_valid; r.cost_avg_valid = p.cost_avg_valid;
      r.residual_block_mean = p.residual_block_mean; r.residual_block_valid_mean = p.residual_block_valid_mean;
    }
    r.trust_region_radius = it.trust_region_radius;
    r.linear_solver_iterations = it.linear_solver_iterations;
    r.iteration_time = it.iteration_time_in_seconds;
    r.cumulative_time = it.cumulative_time_in_seconds;
    r.step_solver_time = it.stage2_time_in_seconds + it.solve_reduced_system_time_in_seconds + it.back_substitution_time_in_seconds;
    r.residual_evaluation_time = it.residual_evaluation_time_in_seconds;
    r.stage1_time = it.stage1_time_in_seconds;
    r.compute_preconditioner_time = it.compute_preconditioner_time_in_seconds;
    r.stage2_time = it.stage2_time_in_seconds;
    r.solve_reduced_system_time = it.solve_reduced_system_time_in_seconds;
    r.back_substitution_time = it.back_substitution_time_in_seconds;
    r.update_cameras_time = it.update_cameras_time_in_seconds;
    rows.push_back(r);
  }
  return rows;
}

inline bool write_ba_log(const std::string& path, const DatasetSummary& ds, const PipelineTimingSummary& pt,
                         const SolverSummary& summary, const std::string& solver_type = "bal_qr") {
  using detail::column; using detail::json_num; using detail::json_str;
  std::ofstream f(path);
  if (!f.is_open()) return false;
  const std::vector<BaIterationRow> rows = make_ba_iterations(summary);
  typedef const BaIterationRow& R;
  const auto B = [](bool v) { return std::string(v ? "true" : "false"); };
  const auto zero = [](R) { return std::string("0"); };
  const auto zerof = [](R) { return std::string("0.0"); };
  f << "{\n";
  // keys in the order of BaLog::BaIteration (ba_log.hpp:139-237)
  column(f, "iteration", rows, [](R r) { return std::to_string(r.iteration); });
  column(f, "linear_solver_type", rows, [](R r) { return json_str(r.linear_solver_type); });
  column(f, "step_is_valid", rows, [&](R r) { return B(r.step_is_valid); });
  column(f, "step_is_nonmonotonic", rows, [&](R r) { return B(r.step_is_nonmonotonic); });
  column(f, "step_is_successful", rows, [&](R r) { return B(r.step_is_successful); });
  column(f, "num_obs", rows, [](R r) { return std::to_string(r.num_obs); });
  column(f, "num_obs_valid", rows, [](R r) { return std::to_string(r.num_obs_valid); });
  column(f, "num_obs_valid_change", rows, [](R r) { return std::to_string(r.num_obs_valid_change); });
  column(f, "cost", rows, [](R r) { return json_num(r.cost); });
  column(f, "cost_change", rows, [](R r) { return json_num(r.cost_change); });
  column(f, "cost_valid", rows, [](R r) { return json_num(r.cost_valid); });
  column(f, "cost_valid_change", rows, [](R r) { return json_num(r.cost_valid_change); });
  column(f, "cost_avg_valid", rows, [](R r) { return json_num(r.cost_avg_valid); });
  column(f, "cost_avg_valid_change", rows, [](R r) { return json_num(r.cost_avg_valid_change); });
  column(f, "grad_projected_norm", rows, zerof);      // Ceres only
  column(f, "grad_projected_max_norm", rows, zerof);  // Ceres only
  column(f, "grad_norm", rows, zerof);                // not computed by optimize_lm_ours either
  column(f, "grad_max_norm", rows, zerof);
  column(f, "residual_block_mean", rows, [](R r) { return json_num(r.residual_block_mean); });
  column(f, "residual_block_valid_mean", rows, [](R r) { return json_num(r.residual_block_valid_mean); });
  column(f, "step_norm", rows, zerof);
  column(f, "relative_decrease", rows, [](R r) { return json_num(r.relative_decrease); });
  column(f, "trust_region_radius", rows, [](R r) { return json_num(r.trust_region_radius); });
  column(f, "linear_solver_iterations", rows, [](R r) { return std::to_string(r.linear_solver_iterations); });
  column(f, "iteration_time", rows, [](R r) { return json_num(r.iteration_time); });
  column(f, "cumulative_time", rows, [](R r) { return json_num(r.cumulative_time); });
  column(f, "logging_time", rows, zerof);
  column(f, "step_solver_time", rows, [](R r) { return json_num(r.step_solver_time); });
  column(f, "residual_evaluation_time", rows, [](R r) { return json_num(r.residual_evaluation_time); });
  column(f, "jacobian_evaluation_time", rows, zerof);      // inside stage1_time (staged execution)
  column(f, "scale_landmark_jacobian_time", rows, zerof);  // inside stage1_time
  column(f, "perform_qr_time", rows, zerof);               // inside stage1_time
  column(f, "stage1_time", rows, [](R r) { return json_num(r.stage1_time); });
  column(f, "scale_pose_jacobian_time", rows, zerof);      // folded into stage 1 on the GPU
  column(f, "landmark_damping_time", rows, zerof);         // inside stage2_time
  column(f, "compute_preconditioner_time", rows, [](R r) { return json_num(r.compute_preconditioner_time); });
  column(f, "compute_gradient_time", rows, zerof);         // inside stage2_time
  column(f, "stage2_time", rows, [](R r) { return json_num(r.stage2_time); });
  column(f, "prepare_time", rows, zerof);
  column(f, "solve_reduced_system_time", rows, [](R r) { return json_num(r.solve_reduced_system_time); });
  column(f, "back_substitution_time", rows, [](R r) { return json_num(r.back_substitution_time); });
  column(f, "update_cameras_time", rows, [](R r) { return json_num(r.update_cameras_time); });
  column(f, "resident_memory", rows, zero);
  column(f, "resident_memory_peak", rows, zero);
  f << "    \"_type\": \"rootba\",\n";
  // _static (ba_log.hpp:54-136)
  int n_ok = -1, n_bad = 0;  // finish_solve, bal_bundle_adjustment.cpp:117-127: iteration 0 is not a step
  double lin_t = 0, res_t = 0;
  for (const auto& it : summary.iterations) {
    if (it.step_is_successful) ++n_ok; else ++n_bad;
    lin_t += it.stage2_time_in_seconds + it.solve_reduced_system_time_in_seconds + it.back_substitution_time_in_seconds;
    res_t += it.residual_evaluation_time_in_seconds;
  }
  const auto stats = [&](const DatasetStats& s) {
    return "{\"mean\": " + json_num(s.mean) + ", \"min\": " + json_num(s.min) + ", \"max\": " + json_num(s.max) + ", \"stddev\": " + json_num(s.stddev) + "}";
  };
  f << "    \"_static\": {\n";
  f << "        \"problem_info\": {\"type\": " << json_str(ds.type) << ", \"input_path\": " << json_str(ds.input_path)
    << ", \"num_cameras\": " << ds.num_cameras << ", \"num_landmarks\": " << ds.num_landmarks << ", \"num_observations\": " << ds.num_observations
    << ", \"rcs_sparsity\": " << json_num(ds.rcs_sparsity) << ", \"per_lm_obs\": " << stats(ds.per_lm_obs) << ", \"per_host_lms\": " << stats(ds.per_host_lms) << "},\n";
  f << "        \"timing\": {\"total\": " << json_num(pt.load_time + pt.preprocess_time + pt.optimize_time) << ", \"load\": " << json_num(pt.load_time)
    << ", \"preprocess\": " << json_num(pt.preprocess_time) << ", \"optimize\": " << json_num(pt.optimize_time) << ", \"postprocess\": " << json_num(pt.postprocess_time) << "},\n";
  f << "        \"solver\": {\"solver_type\": " << json_str(solver_type) << ", \"termination_type\": " << termination_type_code(summary.termination_type)
    << ", \"message\": " << json_str(summary.message) << ", \"num_successful_steps\": " << n_ok << ", \"num_unsuccessful_steps\": " << n_bad
    << ", \"logging_time_in_seconds\": 0.0, \"preprocessor_time_in_seconds\": " << json_num(summary.preprocessor_time_in_seconds)
    << ", \"minimizer_time_in_seconds\": " << json_num(summary.minimizer_time_in_seconds) << ", \"postprocessor_time_in_seconds\": 0.0"
    << ", \"total_time_in_seconds\": " << json_num(summary.total_time_in_seconds) << ", \"linear_solver_time_in_seconds\": " << json_num(lin_t)
    << ", \"num_linear_solves\": " << summary.num_linear_solves << ", \"residual_evaluation_time_in_seconds\": " << json_num(res_t)
    << ", \"num_residual_evaluations\": " << summary.num_residual_evaluations << ", \"jacobian_evaluation_time_in_seconds\": 0.0"
    << ", \"num_jacobian_evaluations\": " << summary.num_jacobian_evaluations << ", \"num_threads_given\": 0, \"num_threads_used\": 0"
    << ", \"num_threads_available\": 0, \"resident_memory_peak\": 0}\n";
  f << "    }\n}\n";
  return f.good();
}

}  // namespace rootba_b200

// CPU test driver (TEST INFRASTRUCTURE -- links the oracle, never shipped): runs the C++ host's LM loop
// (rootba_b200::optimize_lm, rootba_b200/host/solver.hpp, mirror of optimize_lm_ours,
// solver/bal_bundle_adjustment.cpp:249-544) with an ORACLE-backed Linearizor instead of the GPU one and prints the
// trajectory with full precision.  tests/test_lm_loop_cpu.py compares it with the oracle's own LM loop on the same
// arrays: identical arithmetic underneath, so every cost, decision and lambda must be equal.
//
//   lm_loop_cpu --input <bal file> --dump <arrays.bin> [--float] [--max-num-iterations N] [--jacobi] [--huber X]
//               [--optimized-cost ERROR|ERROR_VALID|ERROR_VALID_AVG] [--min-relative-decrease X]
//               [--initial-trust-region-radius X] [--min-trust-region-radius X]
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "../../oracle/rootba_oracle.hpp"
#include "../../rootba_b200/host/solver.hpp"

using namespace rootba_b200;

template <class S>
struct OracleLinearizor {  // members of rootba::Linearizor (solver/linearizor.hpp:56-82) + backup / restore / download_state
  orc::Problem<S> P;
  IterationSummary* it = nullptr;
  void start_iteration(IterationSummary* s) { it = s; }
  void compute_error(ResidualInfo& ri) {
    orc::ResidualInfo r;
    P.compute_error(r);
    ri.all = {r.all.num_obs, r.all.error, r.all.residual_sum};
    ri.valid = {r.valid.num_obs, r.valid.error, r.valid.residual_sum};
    ri.is_numerically_valid = r.is_numerically_valid;
  }
  void linearize() { if (!P.linearize()) throw std::runtime_error("did not expect numerical failure during linearization"); }
  std::vector<S> solve(S lambda) {
    std::vector<S> inc;
    P.solve(lambda, inc);
    if (it) it->linear_solver_iterations = P.last_cg_iterations;
    return inc;
  }
  S apply(std::vector<S>&& inc) { return P.apply(inc); }
  void backup() { P.backup(); }
  void restore() { P.restore(); }
  void download_state() {}
};

template <class S>
int run(const std::string& input, const std::string& dump, const SolverOptions& o) {
  auto problem = load_normalized_bal_problem_parallel<S>(input, true, 100.0, 2);
  std::vector<int64_t> off; std::vector<int32_t> oc; std::vector<S> xy, cams, lms;
  problem.export_topology(off, oc, xy);
  problem.export_state(cams, lms);
  {  // the arrays the oracle is built from, as doubles (exact for float too), in the bal_qr --dump-problem layout
    std::ofstream f(dump, std::ios::binary);
    const int64_t hdr[3] = {problem.num_cameras(), problem.num_landmarks(), (int64_t)oc.size()};
    std::vector<double> c(cams.begin(), cams.end()), l(lms.begin(), lms.end()), x(xy.begin(), xy.end());
    f.write((const char*)hdr, sizeof(hdr));
    f.write((const char*)c.data(), c.size() * 8); f.write((const char*)l.data(), l.size() * 8);
    f.write((const char*)off.data(), off.size() * 8); f.write((const char*)oc.data(), oc.size() * 4); f.write((const char*)x.data(), x.size() * 8);
  }
  OracleLinearizor<S> lin;
  orc::Options& po = lin.P.opt;
  po.use_householder = o.use_householder_marginalization;
  po.use_valid_projections_only = o.use_projection_validity_check();
  po.robust_norm = (int)o.robust_norm;
  po.huber_parameter = o.huber_parameter;
  po.jacobi_scaling_epsilon = o.jacobi_scaling_epsilon;
  po.preconditioner_type = (int)o.preconditioner_type;
  po.min_linear_solver_iterations = o.min_linear_solver_iterations;
  po.max_linear_solver_iterations = o.max_linear_solver_iterations;
  po.eta = o.eta;
  po.num_threads = 1;
  lin.P.init(problem.num_cameras(), problem.num_landmarks(), off.data(), oc.data(), xy.data(), cams.data(), lms.data());
  SolverSummary summary;
  optimize_lm<S>(lin, o, summary, /*quiet=*/true);
  std::printf("termination %s\n", summary.termination_type.c_str());
  for (const auto& it : summary.iterations)
    std::printf("it %d cost %.17g cost_valid %.17g ok %d valid %d trr %.17g rho %.17g cg %d\n", it.iteration, it.cost.all.error, it.cost.valid.error,
                (int)it.step_is_successful, (int)it.step_is_valid, it.trust_region_radius, it.relative_decrease, it.linear_solver_iterations);
  return 0;
}

int main(int argc, char** argv) {
  std::string input, dump;
  bool use_float = false;
  SolverOptions o;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    auto next = [&]() -> std::string { if (i + 1 >= argc) { std::cerr << "missing value for " << a << "\n"; std::exit(2); } return argv[++i]; };
    if (a == "--input") input = next();
    else if (a == "--dump") dump = next();
    else if (a == "--float") use_float = true;
    else if (a == "--max-num-iterations") o.max_num_iterations = std::stoi(next());
    else if (a == "--jacobi") o.preconditioner_type = SolverOptions::PreconditionerType::JACOBI;
    else if (a == "--givens") o.use_householder_marginalization = false;
    else if (a == "--huber") { o.robust_norm = SolverOptions::RobustNorm::HUBER; o.huber_parameter = std::stod(next()); }
    else if (a == "--optimized-cost") { const std::string v = next(); o.optimized_cost = v == "ERROR" ? SolverOptions::OptimizedCost::ERROR : v == "ERROR_VALID" ? SolverOptions::OptimizedCost::ERROR_VALID : SolverOptions::OptimizedCost::ERROR_VALID_AVG; }
    else if (a == "--min-relative-decrease") o.min_relative_decrease = std::stod(next());
    else if (a == "--initial-trust-region-radius") o.initial_trust_region_radius = std::stod(next());
    else if (a == "--min-trust-region-radius") o.min_trust_region_radius = std::stod(next());
    else { std::cerr << "unknown option " << a << "\n"; return 2; }
  }
  if (input.empty() || dump.empty()) { std::cerr << "--input and --dump are required\n"; return 2; }
  try {
    return use_float ? run<float>(input, dump, o) : run<double>(input, dump, o);
  } catch (const std::exception& e) {
    std::cerr << "FATAL: " << e.what() << "\n";
    return 1;
  }
}

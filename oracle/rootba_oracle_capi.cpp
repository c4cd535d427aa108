// C API (ctypes-friendly) over the CPU oracle.  TEST INFRASTRUCTURE ONLY -- see the
// header of rootba_oracle.hpp ("parity unpinned").  Built by oracle/Makefile into
// oracle/_build/librootba_oracle.so.
#include "rootba_oracle.hpp"

using namespace orc;

namespace {

template <class S>
struct Handle {
  Problem<S> P;
  std::vector<typename Problem<S>::SCBlock> sc;
  std::vector<S> sc_diag2;
  std::vector<IterationLog> log;
};

struct COptions {  // mirrored by tests/_oracle.py (keep in sync)
  int use_householder;
  int use_valid_projections_only;
  int robust_norm;
  double huber_parameter;
  double jacobi_scaling_epsilon;
  int preconditioner_type;
  int min_linear_solver_iterations;
  int max_linear_solver_iterations;
  double eta;
  int staged_execution;
  int reduction_alg;
  int max_num_iterations;
  double initial_trust_region_radius;
  double min_trust_region_radius;
  double max_trust_region_radius;
  double min_relative_decrease;
  double function_tolerance;
  double initial_vee;
  double vee_factor;
  int num_threads;
  int optimized_cost;
  int verbose;
};

Options to_options(const COptions* c) {
  Options o;
  if (!c) return o;
  o.use_householder = c->use_householder;
  o.use_valid_projections_only = c->use_valid_projections_only;
  o.robust_norm = c->robust_norm;
  o.huber_parameter = c->huber_parameter;
  o.jacobi_scaling_epsilon = c->jacobi_scaling_epsilon;
  o.preconditioner_type = c->preconditioner_type;
  o.min_linear_solver_iterations = c->min_linear_solver_iterations;
  o.max_linear_solver_iterations = c->max_linear_solver_iterations;
  o.eta = c->eta;
  o.staged_execution = c->staged_execution;
  o.reduction_alg = c->reduction_alg;
  o.max_num_iterations = c->max_num_iterations;
  o.initial_trust_region_radius = c->initial_trust_region_radius;
  o.min_trust_region_radius = c->min_trust_region_radius;
  o.max_trust_region_radius = c->max_trust_region_radius;
  o.min_relative_decrease = c->min_relative_decrease;
  o.function_tolerance = c->function_tolerance;
  o.initial_vee = c->initial_vee;
  o.vee_factor = c->vee_factor;
  o.num_threads = c->num_threads;
  o.optimized_cost = c->optimized_cost;
  o.verbose = c->verbose;
  return o;
}

BalData g_bal;  // staging area for orc_bal_load -> orc_bal_get

}  // namespace

extern "C" {

void orc_default_options(COptions* c) {
  Options o;
  c->use_householder = o.use_householder;
  c->use_valid_projections_only = o.use_valid_projections_only;
  c->robust_norm = o.robust_norm;
  c->huber_parameter = o.huber_parameter;
  c->jacobi_scaling_epsilon = o.jacobi_scaling_epsilon;
  c->preconditioner_type = o.preconditioner_type;
  c->min_linear_solver_iterations = o.min_linear_solver_iterations;
  c->max_linear_solver_iterations = o.max_linear_solver_iterations;
  c->eta = o.eta;
  c->staged_execution = o.staged_execution;
  c->reduction_alg = o.reduction_alg;
  c->max_num_iterations = o.max_num_iterations;
  c->initial_trust_region_radius = o.initial_trust_region_radius;
  c->min_trust_region_radius = o.min_trust_region_radius;
  c->max_trust_region_radius = o.max_trust_region_radius;
  c->min_relative_decrease = o.min_relative_decrease;
  c->function_tolerance = o.function_tolerance;
  c->initial_vee = o.initial_vee;
  c->vee_factor = o.vee_factor;
  c->num_threads = o.num_threads;
  c->optimized_cost = o.optimized_cost;
  c->verbose = o.verbose;
}

int orc_max_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// ---- BAL loader (double) ----
int orc_bal_load(const char* path, int do_normalize, double scale, int* nc, int* nl, int64_t* nobs) {
  g_bal = BalData();
  // autodetect_input_type (bal_problem.cpp:124-135): "bundle" in the file name -> bundler format, else BAL
  std::string name(path);
  const size_t slash = name.find_last_of('/');
  if (slash != std::string::npos) name = name.substr(slash + 1);
  int rc = name.find("bundle") != std::string::npos ? load_bundler(path, g_bal) : load_bal(path, g_bal);
  if (rc != 0) return rc;
  if (do_normalize) normalize(g_bal, scale);
  *nc = g_bal.nc; *nl = g_bal.nl; *nobs = g_bal.nobs;
  return 0;
}
// filter_obs on the staged problem (after normalisation, like load_normalized_bal_problem); returns the new sizes
void orc_bal_filter_obs(double threshold, int* nl, int64_t* nobs) {
  filter_obs(g_bal, threshold);
  *nl = g_bal.nl; *nobs = g_bal.nobs;
}
void orc_bal_get(double* cams, double* lms, int64_t* lm_off, int* obs_cam, double* obs_xy) {
  std::memcpy(cams, g_bal.cams.data(), g_bal.cams.size() * sizeof(double));
  std::memcpy(lms, g_bal.lms.data(), g_bal.lms.size() * sizeof(double));
  std::memcpy(lm_off, g_bal.lm_off.data(), g_bal.lm_off.size() * sizeof(int64_t));
  std::memcpy(obs_cam, g_bal.obs_cam.data(), g_bal.obs_cam.size() * sizeof(int));
  std::memcpy(obs_xy, g_bal.obs_xy.data(), g_bal.obs_xy.size() * sizeof(double));
}
// normalise arrays in place (double), same arithmetic as the loader path
void orc_normalize(int nc, int nl, double* cams, double* lms, double scale) {
  BalData D;
  D.nc = nc; D.nl = nl;
  D.cams.assign(cams, cams + (size_t)10 * nc);
  D.lms.assign(lms, lms + (size_t)3 * nl);
  normalize(D, scale);
  std::memcpy(cams, D.cams.data(), D.cams.size() * sizeof(double));
  std::memcpy(lms, D.lms.data(), D.lms.size() * sizeof(double));
}

#define ORC_API(SFX, S)                                                                                   \
  void* orc_create_##SFX(int nc, int nl, const int64_t* lm_off, const int* obs_cam, const S* obs_xy,     \
                         const S* cams, const S* lms, const COptions* opt) {                             \
    auto* h = new Handle<S>();                                                                            \
    h->P.opt = to_options(opt);                                                                           \
    h->P.init(nc, nl, lm_off, obs_cam, obs_xy, cams, lms);                                                \
    return h;                                                                                             \
  }                                                                                                       \
  void orc_destroy_##SFX(void* hv) { delete (Handle<S>*)hv; }                                             \
  void orc_set_options_##SFX(void* hv, const COptions* opt) { ((Handle<S>*)hv)->P.opt = to_options(opt); } \
  void orc_get_state_##SFX(void* hv, S* cams, S* lms) {                                                   \
    auto& P = ((Handle<S>*)hv)->P;                                                                        \
    std::memcpy(cams, P.cams.data(), P.cams.size() * sizeof(S));                                          \
    std::memcpy(lms, P.lms.data(), P.lms.size() * sizeof(S));                                             \
  }                                                                                                       \
  void orc_set_state_##SFX(void* hv, const S* cams, const S* lms) {                                       \
    auto& P = ((Handle<S>*)hv)->P;                                                                        \
    std::memcpy(P.cams.data(), cams, P.cams.size() * sizeof(S));                                          \
    std::memcpy(P.lms.data(), lms, P.lms.size() * sizeof(S));                                             \
  }                                                                                                       \
  void orc_backup_##SFX(void* hv) { ((Handle<S>*)hv)->P.backup(); }                                       \
  void orc_restore_##SFX(void* hv) { ((Handle<S>*)hv)->P.restore(); }                                     \
  /* out: [all.num_obs, all.error, all.residual_sum, valid.num_obs, valid.error, valid.residual_sum, ok] */ \
  void orc_compute_error_##SFX(void* hv, double* out7) {                                                  \
    ResidualInfo ri;                                                                                      \
    ((Handle<S>*)hv)->P.compute_error(ri);                                                                \
    out7[0] = ri.all.num_obs; out7[1] = ri.all.error; out7[2] = ri.all.residual_sum;                      \
    out7[3] = ri.valid.num_obs; out7[4] = ri.valid.error; out7[5] = ri.valid.residual_sum;                \
    out7[6] = ri.is_numerically_valid ? 1.0 : 0.0;                                                        \
  }                                                                                                       \
  /* low level LinearizationQR API */                                                                     \
  int orc_stage1_##SFX(void* hv, S* diag2_out, int jacobi_blocks) {                                       \
    auto& P = ((Handle<S>*)hv)->P;                                                                        \
    std::vector<S> d;                                                                                     \
    bool ok = P.get_stage1(d, jacobi_blocks != 0);                                                        \
    std::memcpy(diag2_out, d.data(), d.size() * sizeof(S));                                               \
    return ok ? 0 : 1;                                                                                    \
  }                                                                                                       \
  void orc_set_pose_damping_##SFX(void* hv, S lambda) { ((Handle<S>*)hv)->P.set_pose_damping(lambda); }   \
  void orc_stage2_##SFX(void* hv, S lambda, const S* scaling_or_null, int schur_blocks, S* b_out,         \
                        S* blocks_out_or_null) {                                                          \
    auto& P = ((Handle<S>*)hv)->P;                                                                        \
    std::vector<S> b;                                                                                     \
    P.get_stage2(lambda, scaling_or_null, schur_blocks != 0, b);                                          \
    std::memcpy(b_out, b.data(), b.size() * sizeof(S));                                                   \
    if (blocks_out_or_null)                                                                               \
      std::memcpy(blocks_out_or_null, P.precond_blocks.data(), P.precond_blocks.size() * sizeof(S));      \
  }                                                                                                       \
  void orc_right_multiply_##SFX(void* hv, const S* x, S* y) { ((Handle<S>*)hv)->P.right_multiply(x, y); } \
  int orc_back_substitute_##SFX(void* hv, const S* pose_inc, S* l_diff) {                                 \
    bool fail = false;                                                                                    \
    *l_diff = ((Handle<S>*)hv)->P.back_substitute(pose_inc, fail);                                        \
    return fail ? 1 : 0;                                                                                  \
  }                                                                                                       \
  void orc_block_dims_##SFX(void* hv, int lm, int* rows, int* cols, int* lm_idx, int* res_idx) {          \
    auto& b = ((Handle<S>*)hv)->P.blocks[lm];                                                             \
    *rows = b.num_rows; *cols = b.num_cols; *lm_idx = b.lm_idx; *res_idx = b.res_idx;                     \
  }                                                                                                       \
  void orc_get_block_##SFX(void* hv, int lm, S* out, S* jl_col_scale3) {                                  \
    auto& b = ((Handle<S>*)hv)->P.blocks[lm];                                                             \
    std::memcpy(out, b.storage.data(), b.storage.size() * sizeof(S));                                     \
    if (jl_col_scale3) for (int d = 0; d < 3; ++d) jl_col_scale3[d] = b.Jl_col_scale[d];                  \
  }                                                                                                       \
  /* LinearizorQR API */                                                                                  \
  int orc_linearize_##SFX(void* hv) { return ((Handle<S>*)hv)->P.linearize() ? 0 : 1; }                   \
  void orc_get_scaling_##SFX(void* hv, S* out) {                                                          \
    auto& P = ((Handle<S>*)hv)->P;                                                                        \
    std::memcpy(out, P.pose_jacobian_scaling.data(), P.pose_jacobian_scaling.size() * sizeof(S));        \
  }                                                                                                       \
  int orc_solve_##SFX(void* hv, S lambda, S* inc_out, S* b_out_or_null, S* inv_out_or_null,               \
                      int* cg_iterations, int* cg_termination) {                                          \
    auto& P = ((Handle<S>*)hv)->P;                                                                        \
    std::vector<S> inc, b, inv;                                                                           \
    P.solve(lambda, inc, &b, &inv);                                                                       \
    std::memcpy(inc_out, inc.data(), inc.size() * sizeof(S));                                             \
    if (b_out_or_null) std::memcpy(b_out_or_null, b.data(), b.size() * sizeof(S));                        \
    if (inv_out_or_null) std::memcpy(inv_out_or_null, inv.data(), inv.size() * sizeof(S));                \
    if (cg_iterations) *cg_iterations = P.last_cg_iterations;                                             \
    if (cg_termination) *cg_termination = P.last_cg_termination;                                          \
    return 0;                                                                                             \
  }                                                                                                       \
  void orc_apply_##SFX(void* hv, const S* inc, S* l_diff) {                                               \
    auto& P = ((Handle<S>*)hv)->P;                                                                        \
    std::vector<S> v(inc, inc + (size_t)9 * P.nc);                                                        \
    *l_diff = P.apply(v);                                                                                 \
  }                                                                                                       \
  /* timings of the last calls: stage1, stage2, precond, pcg, backsub, update, error ; + matvec count */  \
  void orc_get_timings_##SFX(void* hv, double* t8) {                                                      \
    auto& P = ((Handle<S>*)hv)->P;                                                                        \
    t8[0] = P.t_stage1; t8[1] = P.t_stage2; t8[2] = P.t_precond; t8[3] = P.t_pcg;                         \
    t8[4] = P.t_backsub; t8[5] = P.t_update; t8[6] = P.t_error; t8[7] = (double)P.total_matvecs;          \
  }                                                                                                       \
  /* LM loop ; per-iteration log rows of 20 doubles ; returns number of rows, termination via *term */    \
  int orc_optimize_##SFX(void* hv, double* log_out, int max_rows, int* term) {                            \
    auto* h = (Handle<S>*)hv;                                                                             \
    int t = optimize_lm<S>(h->P, h->log);                                                                 \
    if (term) *term = t;                                                                                  \
    int n = (int)std::min<size_t>(h->log.size(), (size_t)max_rows);                                       \
    for (int i = 0; i < n; ++i) {                                                                         \
      const auto& L = h->log[i];                                                                          \
      double* r = log_out + (size_t)20 * i;                                                               \
      r[0] = L.iteration; r[1] = L.cost; r[2] = L.cost_valid; r[3] = L.num_obs_valid;                     \
      r[4] = L.step_is_valid; r[5] = L.step_is_successful; r[6] = L.lambda; r[7] = L.trust_region_radius; \
      r[8] = L.relative_decrease; r[9] = L.l_diff; r[10] = L.cg_iterations; r[11] = L.stage1_time;        \
      r[12] = L.stage2_time; r[13] = L.precond_time; r[14] = L.pcg_time; r[15] = L.backsub_time;          \
      r[16] = L.update_time; r[17] = L.error_time; r[18] = L.iteration_time; r[19] = 0;                   \
    }                                                                                                     \
    return n;                                                                                             \
  }                                                                                                       \
  /* Schur-complement cross-check (reference's own QR validation) */                                      \
  void orc_sc_linearize_##SFX(void* hv, S* diag2_out) {                                                   \
    auto* h = (Handle<S>*)hv;                                                                             \
    h->P.sc_linearize(h->sc, h->sc_diag2);                                                                \
    if (diag2_out) std::memcpy(diag2_out, h->sc_diag2.data(), h->sc_diag2.size() * sizeof(S));           \
  }                                                                                                       \
  void orc_sc_scale_Jp_##SFX(void* hv, const S* scaling) { auto* h = (Handle<S>*)hv; h->P.sc_scale_Jp(h->sc, scaling); } \
  void orc_sc_get_Hb_##SFX(void* hv, S lambda, S pose_damping, const S* x, S* b_out, S* diag_blocks_out, \
                           S* y_out) {                                                                    \
    auto* h = (Handle<S>*)hv;                                                                             \
    std::vector<S> b, d, y;                                                                               \
    h->P.sc_get_Hb(h->sc, lambda, pose_damping, b, d, x, y);                                              \
    std::memcpy(b_out, b.data(), b.size() * sizeof(S));                                                   \
    std::memcpy(diag_blocks_out, d.data(), d.size() * sizeof(S));                                         \
    std::memcpy(y_out, y.data(), y.size() * sizeof(S));                                                   \
  }                                                                                                       \
  void orc_scl_linearize_##SFX(void* hv) { ((Handle<S>*)hv)->P.sc_linearizor_linearize(); }                      \
  void orc_scl_get_scaling_##SFX(void* hv, S* out) { auto* h = (Handle<S>*)hv; std::memcpy(out, h->P.sc_scaling.data(), h->P.sc_scaling.size() * sizeof(S)); } \
  void orc_scl_solve_##SFX(void* hv, S lambda, S* inc, S* b_out, S* inv_out, int* it, int* term) {              \
    auto* h = (Handle<S>*)hv;                                                                                 \
    std::vector<S> v, b, inv;                                                                                 \
    h->P.sc_linearizor_solve(lambda, v, &b, &inv);                                                            \
    std::memcpy(inc, v.data(), v.size() * sizeof(S));                                                         \
    if (b_out) std::memcpy(b_out, b.data(), b.size() * sizeof(S));                                            \
    if (inv_out) std::memcpy(inv_out, inv.data(), inv.size() * sizeof(S));                                    \
    *it = h->P.last_cg_iterations; *term = h->P.last_cg_termination;                                          \
  }                                                                                                           \
  void orc_scl_power_solve_##SFX(void* hv, S lambda, int order, S q_tol, S* inc, S* b_out, int* it, int* term) { \
    auto* h = (Handle<S>*)hv;                                                                                 \
    std::vector<S> v, b;                                                                                      \
    h->P.power_sc_linearizor_solve(lambda, order, q_tol, v, &b);                                              \
    std::memcpy(inc, v.data(), v.size() * sizeof(S));                                                         \
    if (b_out) std::memcpy(b_out, b.data(), b.size() * sizeof(S));                                            \
    *it = h->P.last_power_order; *term = h->P.last_power_termination;                                         \
  }                                                                                                           \
  void orc_scl_e0_##SFX(void* hv, S lambda, const S* x, S* y) { ((Handle<S>*)hv)->P.sc_right_mul_e0(lambda, x, y); } \
  void orc_scl_apply_##SFX(void* hv, const S* inc, S* l_diff) {                                               \
    auto* h = (Handle<S>*)hv;                                                                                 \
    std::vector<S> v(inc, inc + (size_t)9 * h->P.nc);                                                         \
    *l_diff = h->P.sc_linearizor_apply(v);                                                                    \
  }                                                                                                           \
  void orc_sc_back_substitute_##SFX(void* hv, S lambda, const S* pose_inc, S* l_diff, S* lms_out) {       \
    auto* h = (Handle<S>*)hv;                                                                             \
    std::vector<S> l;                                                                                     \
    *l_diff = h->P.sc_back_substitute(h->sc, lambda, pose_inc, l);                                        \
    std::memcpy(lms_out, l.data(), l.size() * sizeof(S));                                                 \
  }                                                                                                       \
  /* single-observation helpers for the Jacobian / projection property tests */                          \
  int orc_linearize_point_##SFX(const S* obs, const S* p_w, const S* cam10, S* res2, S* Jp12, S* Ji6,     \
                                S* Jl6) {                                                                 \
    return linearize_point<S>(obs, p_w, cam10, true, res2, Jp12, Ji6, Jl6) ? 1 : 0;                       \
  }                                                                                                       \
  void orc_camera_apply_inc_##SFX(S* cam10, const S* inc9) { camera_apply_inc<S>(cam10, inc9); }          \
  int orc_invert_block9_##SFX(const S* in81, const S* diag9_or_null, S* out81) {                          \
    return Problem<S>::invert_block9(in81, diag9_or_null, out81) ? 0 : 1;                                 \
  }

ORC_API(f32, float)
ORC_API(f64, double)

}  // extern "C"

// =============================================================================
// TEST INFRASTRUCTURE — NOT PRODUCT CODE.  Nothing under cuopt_b200/ includes,
// links or loads this file; only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / `--impl reference` leg use it (through oracle/pdlp_oracle.py).
//
// CPU restatement (fp64, sequential summation order) of the reference's PDLP
// as NVIDIA/cuopt 25.08 runs it on the GPU.  The reference's own GPU
// implementation cannot be built offline (RAFT/RMM/rapids-cmake are fetched
// from the network), and its SpMV / dot / nrm2 arithmetic executes inside
// closed-source cuSPARSE (cusparseSpMV, CUSPARSE_SPMV_CSR_ALG2) and cuBLAS
// (cublasDdot / cublasDnrm2) of the CUDA 12.9 toolkit, so this file restates
// the published algorithm at the reference's call sites.  Each function cites
// the reference lines it follows (paths relative to
// /root/reference/cpp/src/linear_programming unless noted).
//
// PINNED against the reference's own known answers (tests/test_oracle_pins.py):
//   * afiro Methodical1 initial step size 1.4893 / primal weight 0.0141652 +-1e-4
//       (cpp/tests/linear_programming/pdlp_test.cu:237-283)
//   * afiro default-settings primal vector, 32 values, rel 1e-4
//       (python/cuopt/cuopt/tests/linear_programming/test_lp_solver.py:430-476)
//   * afiro objective -464.7531 (rel 1e-6) (test_lp_solver.py:101-121, :592-606)
//   * good-max 17.0, max_offset 0.0 +-1e-4 (pdlp_test.cu:909-943); ranged C-API LP 32.0
//       (c_api_tests/c_api_test.c:761-874)
//   * objectives of the reference's CPU dual simplex (oracle/_ref) on every on-disk instance
// NOT pinned by any reference test: absolute iteration counts.
//
// Summation order: row sums of the SpMVs run left to right in CSR order; vector
// reductions are chunked (4096 elements per chunk, chunks added left to right),
// which makes results independent of the OpenMP thread count.
// Restart strategies: 1 = KKT (Stable1 / Stable2 / Fast1) and 2 = trust region (Methodical1:
// pdlp_restart_strategy.cu:278-364, 842-1080, 1291-1678, 1681-1900 + utils.cuh:240-345), the
// latter restated sequentially: a stable sort replaces thrust::sort_by_key (order among equal
// thresholds only changes rounding of two sums), and gap_reduction_ratio_last_trial, which the
// reference never initialises (pdlp_restart_strategy.cu:160), starts at 1 as in PDLP.jl.
// =============================================================================
#ifdef _OPENMP
#include <omp.h>
#endif
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

namespace {

constexpr double kInf = std::numeric_limits<double>::infinity();

struct hyper_t {  // pdlp_hyper_params.cu:22-80, field order is the ctypes ABI (oracle/pdlp_oracle.py)
  double initial_step_size_scaling;
  int l_inf_ruiz_iterations;
  int do_pock_chambolle_scaling;
  int do_ruiz_scaling;
  double alpha_pock_chambolle;
  double artificial_restart_threshold;
  int compute_initial_step_size_before_scaling;
  int compute_initial_primal_weight_before_scaling;
  double initial_primal_weight_c_scaling;
  double initial_primal_weight_b_scaling;
  int major_iteration;
  int min_iteration_restart;
  int restart_strategy;
  int never_restart_to_average;
  double reduction_exponent;
  double growth_exponent;
  double primal_weight_update_smoothing;
  double sufficient_reduction_for_restart;
  double necessary_reduction_for_restart;
  double primal_importance;
  double primal_distance_smoothing;
  double dual_distance_smoothing;
  int compute_last_restart_before_new_primal_weight;
  int artificial_restart_in_main_loop;
  int rescale_for_restart;
  int handle_some_primal_gradients_on_finite_bounds_as_residuals;
  int project_initial_primal;
};

struct settings_t {
  double abs_dual_tol, rel_dual_tol, abs_primal_tol, rel_primal_tol, abs_gap_tol, rel_gap_tol;
  int iteration_limit;
  double time_limit;
  int num_threads;
  int per_constraint_residual;  // convergence_information.cu:163-204, termination_strategy.cu:141-166
  int detect_infeasibility;     // infeasibility_information.cu, termination_strategy.cu:229-250, pdlp.cu:716-770
  int strict_infeasibility;
  double primal_infeasible_tol, dual_infeasible_tol;
  int save_best_primal_so_far;  // pdlp.cu:333-463, :265-331
};

struct stats_t {
  int termination_status;  // constants.h:62-72
  int number_of_steps_taken;
  int total_number_of_attempted_steps;
  double primal_objective, dual_objective, gap, relative_gap;
  double l2_primal_residual, l2_relative_primal_residual;
  double l2_dual_residual, l2_relative_dual_residual;
  double step_size, primal_weight;
  int n_restarts, n_major;
  double solve_seconds;
  int solution_is_average;
};

struct csr_t {
  int rows = 0, cols = 0;
  std::vector<int> off, idx;
  std::vector<double> val;
};

// y = A x, each row summed left to right (what a one-thread-per-row CSR kernel does).
void spmv(const csr_t& A, const double* x, double* y)
{
#pragma omp parallel for schedule(static) if (A.rows > 20000)
  for (int i = 0; i < A.rows; ++i) {
    double s = 0.0;
    for (int p = A.off[i]; p < A.off[i + 1]; ++p)
      s += A.val[p] * x[A.idx[p]];
    y[i] = s;
  }
}

// Chunked deterministic reduction of f(i) over [0,n).
template <typename F>
double reduce_sum(int n, F f)
{
  constexpr int CH = 4096;
  const int nch    = (n + CH - 1) / CH;
  std::vector<double> part(std::max(nch, 1), 0.0);
#pragma omp parallel for schedule(static) if (nch > 8)
  for (int c = 0; c < nch; ++c) {
    double s     = 0.0;
    const int hi = std::min(n, (c + 1) * CH);
    for (int i = c * CH; i < hi; ++i)
      s += f(i);
    part[c] = s;
  }
  double s = 0.0;
  for (int c = 0; c < nch; ++c)
    s += part[c];
  return s;
}

// CSR transpose with stable (row-ascending) order inside each transposed row, the order
// cusparseCsr2cscEx2 produces (reference: mip/problem/problem.cu:277-309).
csr_t transpose(const csr_t& A)
{
  csr_t T;
  T.rows = A.cols;
  T.cols = A.rows;
  T.off.assign(A.cols + 1, 0);
  T.idx.resize(A.idx.size());
  T.val.resize(A.val.size());
  for (int j : A.idx)
    T.off[j + 1]++;
  for (int j = 0; j < A.cols; ++j)
    T.off[j + 1] += T.off[j];
  std::vector<int> cur(T.off.begin(), T.off.end() - 1);
  for (int i = 0; i < A.rows; ++i)
    for (int p = A.off[i]; p < A.off[i + 1]; ++p) {
      const int q = cur[A.idx[p]]++;
      T.idx[q]    = i;
      T.val[q]    = A.val[p];
    }
  return T;
}

// utils.cuh:140-148
inline double combine_finite_abs_bounds(double lower, double upper)
{
  double v = 0.0;
  if (std::isfinite(upper)) v = std::max(v, std::fabs(upper));
  if (std::isfinite(lower)) v = std::max(v, std::fabs(lower));
  return v;
}
// utils.cuh:166-178
inline double violation(double value, double lower, double upper)
{
  if (value < lower) return lower - value;
  if (value > upper) return value - upper;
  return 0.0;
}
// utils.cuh:205-219
inline double bound_value_reduced_cost_product(double value, double lower, double upper)
{
  double bound_value = 0.0;
  if (value > 0.0)
    bound_value = lower;
  else if (value < 0.0)
    bound_value = upper;
  return std::isfinite(bound_value) ? value * bound_value : 0.0;
}
// raft::linalg::eltwiseDivideCheckZero: a / b, 0 when b == 0
inline double div_check_zero(double a, double b) { return b == 0.0 ? 0.0 : a / b; }

struct convergence_t {  // termination_strategy/convergence_information.cu
  double l2_primal_residual = 0, l2_dual_residual = 0, primal_objective = 0, dual_objective = 0;
  double linf_relative_primal_residual = 0, linf_relative_dual_residual = 0;  // per_constraint_residual only
  double max_primal_ray_infeasibility = 0, primal_ray_linear_objective = 0;   // detect_infeasibility only
  double max_dual_ray_infeasibility = 0, dual_ray_linear_objective = 0;
  double gap = 0, abs_objective = 0, l2_primal_variable = 0, l2_dual_variable = 0;
  std::vector<double> reduced_cost;
  int status = 6;
};

struct trace_row_t {
  int k, restarted, to_average;
  double primal_objective, dual_objective, gap, l2_primal_residual, l2_dual_residual, step_size, primal_weight,
    kkt_current, kkt_average;
};

class oracle_t {
 public:
  // ---- unscaled problem after problem_t construction (mip/problem/problem.cu:55-93) ----
  int m, n;
  csr_t A, AT;
  std::vector<double> c, l, u, lc, uc, b_comb;
  double obj_scale, obj_offset;  // objective = obj_scale * c'x + obj_offset  (c already negated when maximising)
  bool maximize;
  hyper_t hp;
  settings_t st;

  // ---- scaled problem (initial_scaling.cu) ----
  csr_t As, ATs;
  std::vector<double> cs, ls, us, lcs, ucs;
  std::vector<double> Dr, Dc;  // cumulative constraint / variable scaling

  // ---- iterate state (saddle_point.cu:26-60: x = y = 0) ----
  std::vector<double> x, y, xn, yn, xbar, dx, dy, AtY, AtYn, Ax;
  std::vector<double> sum_x, sum_y, x_avg, y_avg, x_lr, y_lr, tmp_n, tmp_m;
  double sum_w             = 0.0;
  double step_size         = 0.0, primal_weight = 0.0, tau = 0.0, sigma = 0.0;
  double interaction = 0, norm_dx2 = 0, norm_dy2 = 0;
  int k_total = 0, k_internal = 0, k_pdhg_host = 0, k_pdhg_dev = 0, its_since_restart = 0;
  bool last_restart_was_average = false;
  int valid_step_size           = 0;
  double last_candidate_kkt = 0.0, last_restart_kkt = 0.0;
  double l2_norm_c = 0.0, l2_norm_b = 0.0;
  convergence_t conv_cur, conv_avg;
  bool initialised = false;
  int n_restarts = 0, n_major = 0;
  // warm start handed in before initialise() (pdlp_warm_start_data.hpp:28-72): 9 vectors in the header's order
  // {current primal, current dual, primal average, dual average, current A^T y, sum primal, sum dual,
  //  last-restart primal, last-restart dual} and 8 scalars {primal weight, step size, total pdlp iterations,
  //  total pdhg iterations, last candidate kkt, last restart kkt, sum of solution weights, iterations since restart}
  double last_linf[2] = {0, 0};  // per-constraint residuals of the last pdlp_oracle_convergence call
  double gap_reduction_ratio_last_trial = 1.0;  // trust-region restart (see header)
  bool warm_given = false;
  std::vector<double> warm_v[9];
  double warm_s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  std::vector<trace_row_t> trace;
  stats_t result{};
  std::vector<double> sol_x, sol_y, sol_rc;

  oracle_t(int m_, int n_, const int* off, const int* idx, const double* val, const double* c_, const double* l_,
           const double* u_, const double* lc_, const double* uc_, int maximize_, double offset, const hyper_t& h,
           const settings_t& s)
    : m(m_), n(n_), hp(h), st(s)
  {
    A.rows = m;
    A.cols = n;
    A.off.assign(off, off + m + 1);
    A.idx.assign(idx, idx + off[m]);
    A.val.assign(val, val + off[m]);
    c.assign(c_, c_ + n);
    l.assign(l_, l_ + n);
    u.assign(u_, u_ + n);
    lc.assign(lc_, lc_ + m);
    uc.assign(uc_, uc_ + m);
    maximize   = maximize_ != 0;
    obj_scale  = 1.0;
    obj_offset = offset;
    // mip/problem/problem_helpers.cuh:128-142: maximise => negate c, flip the objective scaling factor
    if (maximize) {
      for (auto& v : c)
        v = -v;
      obj_scale = -obj_scale;
    }
    AT = transpose(A);
    b_comb.resize(m);
    for (int i = 0; i < m; ++i)
      b_comb[i] = combine_finite_abs_bounds(lc[i], uc[i]);
    // convergence_information.cu:74-82: norms of the UNSCALED c and combined bounds (cublasDnrm2)
    l2_norm_c = std::sqrt(reduce_sum(n, [&](int j) { return c[j] * c[j]; }));
    l2_norm_b = std::sqrt(reduce_sum(m, [&](int i) { return b_comb[i] * b_comb[i]; }));
  }

  // ------------------------------------------------------------------ scaling
  // initial_scaling.cu:85-163 (Ruiz, L-inf), :177-307 (Pock-Chambolle)
  void compute_scaling_vectors()
  {
    Dr.assign(m, 1.0);
    Dc.assign(n, 1.0);
    std::vector<double> it_r(m), it_c(n);
    auto bounded_div_sqrt = [](double a, double b) { return b > 0.0 ? a / std::sqrt(b) : a; };  // utils.cuh:123-129
    if (hp.do_ruiz_scaling) {
      for (int it = 0; it < hp.l_inf_ruiz_iterations; ++it) {
        std::fill(it_r.begin(), it_r.end(), 0.0);
        std::fill(it_c.begin(), it_c.end(), 0.0);
        for (int i = 0; i < m; ++i)
          for (int p = A.off[i]; p < A.off[i + 1]; ++p) {
            const int j    = A.idx[p];
            const double a = std::fabs((A.val[p] * Dr[i]) * Dc[j]);  // :104-106
            it_r[i]        = std::max(it_r[i], a);
            it_c[j]        = std::max(it_c[j], a);
          }
        for (int i = 0; i < m; ++i)
          Dr[i] = bounded_div_sqrt(Dr[i], it_r[i]);
        for (int j = 0; j < n; ++j)
          Dc[j] = bounded_div_sqrt(Dc[j], it_c[j]);
      }
    }
    if (hp.do_pock_chambolle_scaling) {
      const double alpha = hp.alpha_pock_chambolle;
      for (int i = 0; i < m; ++i) {  // :177-213 rows of A
        double s = 0.0;
        for (int p = A.off[i]; p < A.off[i + 1]; ++p)
          s += std::pow(std::fabs((A.val[p] * Dr[i]) * Dc[A.idx[p]]), alpha);
        it_r[i] = s;
      }
      for (int j = 0; j < n; ++j) {  // :216-252 rows of A^T, same association (a*Dr)*Dc
        double s = 0.0;
        for (int p = AT.off[j]; p < AT.off[j + 1]; ++p)
          s += std::pow(std::fabs((AT.val[p] * Dr[AT.idx[p]]) * Dc[j]), 2.0 - alpha);
        it_c[j] = s;
      }
      for (int i = 0; i < m; ++i)
        Dr[i] = bounded_div_sqrt(Dr[i], it_r[i]);
      for (int j = 0; j < n; ++j)
        Dc[j] = bounded_div_sqrt(Dc[j], it_c[j]);
    }
  }

  // initial_scaling.cu:310-408
  void scale_problem()
  {
    As  = A;
    ATs = AT;
    for (int i = 0; i < m; ++i)
      for (int p = A.off[i]; p < A.off[i + 1]; ++p)
        As.val[p] = A.val[p] * Dr[i] * Dc[A.idx[p]];  // :310-326
    for (int j = 0; j < n; ++j)
      for (int p = AT.off[j]; p < AT.off[j + 1]; ++p)
        ATs.val[p] = AT.val[p] * Dc[j] * Dr[AT.idx[p]];  // :329-345 (note the other association order)
    cs.resize(n); ls.resize(n); us.resize(n); lcs.resize(m); ucs.resize(m);
    for (int j = 0; j < n; ++j) {
      cs[j] = c[j] * Dc[j];
      ls[j] = div_check_zero(l[j], Dc[j]);
      us[j] = div_check_zero(u[j], Dc[j]);
    }
    for (int i = 0; i < m; ++i) {
      lcs[i] = lc[i] * Dr[i];
      ucs[i] = uc[i] * Dr[i];
    }
    scale_solutions(x, y);  // :404-407
  }
  void scale_solutions(std::vector<double>& px, std::vector<double>& py) const  // :411-427
  {
    for (int j = 0; j < n; ++j) px[j] = div_check_zero(px[j], Dc[j]);
    for (int i = 0; i < m; ++i) py[i] = div_check_zero(py[i], Dr[i]);
  }
  void unscale_solutions(std::vector<double>& px, std::vector<double>& py) const  // :456-484
  {
    for (int j = 0; j < n; ++j) px[j] *= Dc[j];
    for (int i = 0; i < m; ++i) py[i] *= Dr[i];
  }

  // pdlp.cu:1225-1258: step = initial_step_size_scaling / max|a_ij| of the matrix `M` currently in op_problem_scaled_
  void compute_initial_step_size(const csr_t& M)
  {
    double mx = 0.0;
    for (double v : M.val)
      mx = std::max(mx, std::fabs(v));
    step_size = div_check_zero(hp.initial_step_size_scaling, mx);
  }
  // pdlp.cu:1261-1309
  void compute_initial_primal_weight(const std::vector<double>& cc, const std::vector<double>& lo,
                                     const std::vector<double>& hi)
  {
    const double bs = hp.initial_primal_weight_b_scaling, csn = hp.initial_primal_weight_c_scaling;
    const double bn = std::sqrt(reduce_sum(m, [&](int i) {
      const double v = combine_finite_abs_bounds(lo[i], hi[i]);
      return v * v * bs;
    }));
    const double cn = std::sqrt(reduce_sum(n, [&](int j) { return cc[j] * cc[j] * csn; }));
    if (bn > 0.0 && cn > 0.0)
      primal_weight = hp.primal_importance * (cn / bn);
    else
      primal_weight = hp.primal_importance;
  }

  // pdlp.cu:984-1056 up to the start of the while loop
  void initialise()
  {
    x.assign(n, 0.0); y.assign(m, 0.0); xn.assign(n, 0.0); yn.assign(m, 0.0); xbar.assign(n, 0.0);
    dx.assign(n, 0.0); dy.assign(m, 0.0); AtY.assign(n, 0.0); AtYn.assign(n, 0.0); Ax.assign(m, 0.0);
    sum_x.assign(n, 0.0); sum_y.assign(m, 0.0); x_avg.assign(n, 0.0); y_avg.assign(m, 0.0);
    x_lr.assign(n, 0.0); y_lr.assign(m, 0.0); tmp_n.assign(n, 0.0); tmp_m.assign(m, 0.0);
    conv_cur.reduced_cost.assign(n, 0.0);
    conv_avg.reduced_cost.assign(n, 0.0);
    compute_scaling_vectors();  // scaling strategy ctor, initial_scaling.cu:36-83
    if (hp.compute_initial_step_size_before_scaling) compute_initial_step_size(A);
    if (hp.compute_initial_primal_weight_before_scaling) compute_initial_primal_weight(c, lc, uc);
    scale_problem();
    if (!hp.compute_initial_step_size_before_scaling) compute_initial_step_size(As);
    if (!hp.compute_initial_primal_weight_before_scaling) compute_initial_primal_weight(cs, lcs, ucs);
    if (warm_given) {  // pdlp.cu:131-181 (state copied in) + :1010-1038 (scalars, initial solution scaled at :963)
      x = warm_v[0]; y = warm_v[1];
      scale_solutions(x, y);
      x_avg = warm_v[2]; y_avg = warm_v[3];  // stay unscaled: the first major iteration uses them as they are
      AtY = warm_v[4]; sum_x = warm_v[5]; sum_y = warm_v[6]; x_lr = warm_v[7]; y_lr = warm_v[8];
      primal_weight      = warm_s[0];
      step_size          = warm_s[1];
      k_total            = (int)warm_s[2];
      k_pdhg_host        = (int)warm_s[3];
      k_pdhg_dev         = (int)warm_s[3];
      last_candidate_kkt = warm_s[4];
      last_restart_kkt   = warm_s[5];
      sum_w              = warm_s[6];
      its_since_restart  = (int)warm_s[7];
    }
    tau   = step_size / primal_weight;  // adaptive_step_size_strategy.cu:348-366
    sigma = step_size * primal_weight;
    if (hp.project_initial_primal) {  // pdlp.cu:1041-1056, clamp = min(max(v, lo), hi) (utils.cuh:131-137)
      for (int j = 0; j < n; ++j) {
        x[j]     = std::min(std::max(x[j], ls[j]), us[j]);
        x_avg[j] = std::min(std::max(x_avg[j], ls[j]), us[j]);
      }
    }
    initialised = true;
  }

  // --------------------------------------------------------------- PDHG step
  // pdhg.cu:137-158 + utils.cuh:81-95
  void primal_projection()
  {
#pragma omp parallel for schedule(static) if (n > 20000)
    for (int j = 0; j < n; ++j) {
      const double gradient = cs[j] - AtY[j];
      double next           = x[j] - (tau * gradient);
      next                  = std::max(std::min(next, us[j]), ls[j]);
      xn[j]                 = next;
      dx[j]                 = next - x[j];
      xbar[j]               = next - x[j] + next;
    }
  }
  // pdhg.cu:73-117 + utils.cuh:98-112
  void dual_projection()
  {
    spmv(As, xbar.data(), Ax.data());
#pragma omp parallel for schedule(static) if (m > 20000)
    for (int i = 0; i < m; ++i) {
      double next      = y[i] - (sigma * Ax[i]);
      const double low = next + sigma * lcs[i];
      const double up  = next + sigma * ucs[i];
      next             = std::max(low, std::min(up, 0.0));
      yn[i]            = next;
      dy[i]            = next - y[i];
    }
  }
  // adaptive_step_size_strategy.cu:232-345
  void interaction_and_movement()
  {
    spmv(ATs, yn.data(), AtYn.data());
    interaction = reduce_sum(n, [&](int j) { return (AtYn[j] - AtY[j]) * dx[j]; });
    norm_dx2    = reduce_sum(n, [&](int j) { return dx[j] * dx[j]; });
    norm_dy2    = reduce_sum(m, [&](int i) { return dy[i] * dy[i]; });
  }
  // adaptive_step_size_strategy.cu:92-188
  void step_sizes_from_movement_and_interaction()
  {
    const double movement =
      hp.primal_distance_smoothing * primal_weight * norm_dx2 + (hp.dual_distance_smoothing / primal_weight) * norm_dy2;
    if (movement <= 0.0 || movement >= 1.0e100) {
      valid_step_size = -1;
      return;
    }
    const double inter = std::fabs(interaction);
    double eta         = step_size;
    k_pdhg_dev += 1;
    const double kcoef = (double)k_pdhg_dev;
    const double limit = inter > 0.0 ? movement / inter : kInf;
    if (eta <= limit) valid_step_size = 1;
    const double cand1 = (1.0 - std::pow(kcoef + 1.0, -hp.reduction_exponent)) * limit;
    const double cand2 = (1.0 + std::pow(kcoef + 1.0, -hp.growth_exponent)) * eta;
    eta                = std::min(cand1, cand2);
    tau                = eta / primal_weight;
    sigma              = eta * primal_weight;
    step_size          = eta;
  }
  // pdlp.cu:1188-1222
  void take_step()
  {
    valid_step_size = 0;
    while (valid_step_size == 0) {
      // pdhg.cu:183-202: A^T y only on the very first step or right after a restart to the average
      if (k_pdhg_host == 0 || (its_since_restart == 0 && last_restart_was_average)) spmv(ATs, y.data(), AtY.data());
      primal_projection();
      dual_projection();
      k_pdhg_host += 1;  // pdhg.cu:234
      interaction_and_movement();
      step_sizes_from_movement_and_interaction();
    }
    // weighted_average_solution.cu:73-110, weight = the step size AFTER its update (pdlp.cu:1216-1219)
    const double w = step_size;
#pragma omp parallel for schedule(static) if (n > 20000)
    for (int j = 0; j < n; ++j) sum_x[j] = sum_x[j] + w * xn[j];
#pragma omp parallel for schedule(static) if (m > 20000)
    for (int i = 0; i < m; ++i) sum_y[i] = sum_y[i] + w * yn[i];
    sum_w += w;
    its_since_restart += 1;
    x.swap(xn);  // pdhg.cu:238-250
    y.swap(yn);
    AtY.swap(AtYn);
  }

  // --------------------------------------------------- termination evaluation
  // convergence_information.cu:150-422 on the UNSCALED problem; px / py are unscaled iterates.
  void compute_convergence_information(const std::vector<double>& px, const std::vector<double>& py, convergence_t& cv)
  {
    // primal residual (:222-250), objective (:262-286)
    spmv(A, px.data(), tmp_m.data());
    cv.l2_primal_residual = std::sqrt(reduce_sum(m, [&](int i) {
      const double v = violation(tmp_m[i], lc[i], uc[i]);
      return v * v;
    }));
    cv.primal_objective = reduce_sum(n, [&](int j) { return px[j] * c[j]; });
    if (obj_scale != 1.0 || obj_offset != 0.0) cv.primal_objective = obj_scale * cv.primal_objective + obj_offset;
    cv.l2_primal_variable = std::sqrt(reduce_sum(n, [&](int j) { return px[j] * px[j]; }));
    // dual residual (:288-322): g = c - A^T y ; reduced cost (:371-398)
    spmv(AT, py.data(), tmp_n.data());
    for (int j = 0; j < n; ++j) {
      const double g = c[j] - tmp_n[j];
      tmp_n[j]       = g;
      // utils.cuh:196-202 (the g>0 && g<0 test there can never hold)
      const double bound_value = g > 0.0 ? l[j] : u[j];
      double rc;
      if (hp.handle_some_primal_gradients_on_finite_bounds_as_residuals) {  // utils.cuh:222-229
        if (g == 0.0) rc = g;
        else if (std::fabs(px[j] - bound_value) <= std::fabs(px[j])) rc = g;
        else rc = 0.0;
      } else {  // utils.cuh:232-239
        if (g == 0.0) rc = g;
        else if (std::isfinite(bound_value)) rc = g;
        else rc = 0.0;
      }
      cv.reduced_cost[j] = rc;
    }
    cv.l2_dual_residual = std::sqrt(reduce_sum(n, [&](int j) {
      const double r = tmp_n[j] - cv.reduced_cost[j];
      return r * r;
    }));
    // dual objective (:324-369, :400-422)
    double dobj = reduce_sum(m, [&](int i) { return bound_value_reduced_cost_product(py[i], lc[i], uc[i]); });
    dobj += reduce_sum(n, [&](int j) { return bound_value_reduced_cost_product(cv.reduced_cost[j], l[j], u[j]); });
    if (obj_scale != 1.0 || obj_offset != 0.0) dobj = obj_scale * dobj + obj_offset;
    cv.dual_objective    = dobj;
    cv.l2_dual_variable  = std::sqrt(reduce_sum(m, [&](int i) { return py[i] * py[i]; }));
    cv.gap               = std::fabs(cv.primal_objective - cv.dual_objective);  // :137-148
    cv.abs_objective     = std::fabs(cv.primal_objective) + std::fabs(cv.dual_objective);
    // termination_strategy.cu:117-250 (per_constraint_residual = false, no infeasibility detection)
    cv.status              = 6;  // "NumericalError" == keep going (:186)
    const bool optimal_gap = cv.gap <= st.abs_gap_tol + st.rel_gap_tol * cv.abs_objective;
    bool primal_feas       = cv.l2_primal_residual <= st.abs_primal_tol + st.rel_primal_tol * l2_norm_b;
    bool dual_feas         = cv.l2_dual_residual <= st.abs_dual_tol + st.rel_dual_tol * l2_norm_c;
    if (st.per_constraint_residual) {
      // linf of (residual_i - rel * rhs_i), reduction seeded with 0 (convergence_information.cu:163-204,
      // utils.cuh:392-404: the dual residual and c enter SIGNED); compared with the ABSOLUTE tolerance only
      // (termination_strategy.cu:141-166)
      double lp = 0.0, ld = 0.0;
      for (int i = 0; i < m; ++i) lp = std::max(lp, violation(tmp_m[i], lc[i], uc[i]) - st.rel_primal_tol * b_comb[i]);
      for (int j = 0; j < n; ++j) ld = std::max(ld, (tmp_n[j] - cv.reduced_cost[j]) - st.rel_dual_tol * c[j]);
      cv.linf_relative_primal_residual = lp;
      cv.linf_relative_dual_residual   = ld;
      primal_feas                      = lp <= st.abs_primal_tol;
      dual_feas                        = ld <= st.abs_dual_tol;
    }
    if (dual_feas && primal_feas && optimal_gap) cv.status = 1;
    else if (primal_feas) cv.status = 7;
    else if (st.detect_infeasibility) {
      infeasibility_information(px, py, cv);
      // termination_strategy.cu:229-250 (2 = Infeasible, 3 = Unbounded in constants.h:62-72)
      if (cv.dual_ray_linear_objective > 0.0 &&
          cv.max_dual_ray_infeasibility / cv.dual_ray_linear_objective <= st.primal_infeasible_tol)
        cv.status = 2;
      else if (cv.primal_ray_linear_objective < 0.0 &&
               cv.max_primal_ray_infeasibility / -cv.primal_ray_linear_objective <= st.dual_infeasible_tol)
        cv.status = 3;
    }
  }

  // infeasibility_information.cu:183-223 (the iterate itself is the ray estimate) on the UNSCALED problem.
  // tmp_m holds A px, tmp_n holds c - A^T py from compute_convergence_information above.
  void infeasibility_information(const std::vector<double>& px, const std::vector<double>& py, convergence_t& cv)
  {
    double xinf = 0.0, yinf = 0.0;
    for (int j = 0; j < n; ++j) xinf = std::max(xinf, std::fabs(px[j]));
    for (int i = 0; i < m; ++i) yinf = std::max(yinf, std::fabs(py[i]));
    const double xinv = xinf != 0.0 ? 1.0 / xinf : 0.0;  // eltwiseDivideCheckZero
    // homogeneous primal residual (:226-248): A x against the bounds with every finite bound replaced by 0
    double hres = 0.0;
    for (int i = 0; i < m; ++i) {
      const double hl = std::isfinite(lc[i]) ? 0.0 : lc[i], hu = std::isfinite(uc[i]) ? 0.0 : uc[i];
      hres = std::max(hres, std::fabs(violation(tmp_m[i], hl, hu)));
    }
    double max_viol = 0.0;  // compute_max_violation :250-268, utils.cuh:181-193
    for (int j = 0; j < n; ++j) {
      if (std::isfinite(l[j])) max_viol = std::max(max_viol, -px[j]);
      if (std::isfinite(u[j])) max_viol = std::max(max_viol, px[j]);
    }
    double pobj = reduce_sum(n, [&](int j) { return px[j] * c[j]; }) * xinv;  // :270-287
    // homogeneous dual residual (:289-313): gradient -A^T y, its reduced costs with the iterate as the ray
    std::vector<double> rc(n);
    double hdres = 0.0, rcinf = 0.0;
    for (int j = 0; j < n; ++j) {
      const double g = (tmp_n[j] - c[j]);  // tmp_n = c - A^T y  =>  -A^T y
      const double bound_value = g > 0.0 ? l[j] : u[j];
      double r;
      if (hp.handle_some_primal_gradients_on_finite_bounds_as_residuals) {
        if (g == 0.0) r = g;
        else if (std::fabs(px[j] - bound_value) <= std::fabs(px[j])) r = g;
        else r = 0.0;
      } else {
        if (g == 0.0) r = g;
        else if (std::isfinite(bound_value)) r = g;
        else r = 0.0;
      }
      rc[j] = r;
      hdres = std::max(hdres, std::fabs(g - r));
      rcinf = std::max(rcinf, std::fabs(r));
    }
    double dobj = reduce_sum(m, [&](int i) { return bound_value_reduced_cost_product(py[i], lc[i], uc[i]); }) +
                  reduce_sum(n, [&](int j) { return bound_value_reduced_cost_product(rc[j], l[j], u[j]); });
    // compute_remaining_stats_kernel :118-181
    const double scaling = std::max(yinf, rcinf);
    if (scaling != 0.0) { hdres /= scaling; dobj /= scaling; } else { hdres = 0.0; dobj = 0.0; }
    double max_primal;
    if (xinf > 0.0) max_primal = std::max(hres, max_viol) / xinf;
    else { max_primal = 0.0; pobj = 0.0; }
    cv.max_primal_ray_infeasibility = max_primal;
    cv.primal_ray_linear_objective  = pobj;
    cv.max_dual_ray_infeasibility   = hdres;
    cv.dual_ray_linear_objective    = dobj;
  }

  // pdlp_restart_strategy.cu:367-405
  double kkt_score(const convergence_t& cv) const
  {
    const double w2 = primal_weight * primal_weight;
    return std::sqrt(w2 * cv.l2_primal_residual * cv.l2_primal_residual +
                     cv.l2_dual_residual * cv.l2_dual_residual / w2 + cv.gap * cv.gap);
  }

  void fill_solution(const std::vector<double>& px, const std::vector<double>& py, const convergence_t& cv, int status,
                     bool is_avg)
  {
    // termination_strategy.cu:270-357
    sol_x  = px;
    sol_y  = py;
    sol_rc = cv.reduced_cost;
    result.termination_status              = status;
    result.number_of_steps_taken           = k_internal;
    result.total_number_of_attempted_steps = k_pdhg_host;
    result.primal_objective                = cv.primal_objective;
    result.dual_objective                  = cv.dual_objective;
    result.gap                             = cv.gap;
    result.relative_gap                    = cv.gap / (1.0 + std::fabs(cv.primal_objective) + std::fabs(cv.dual_objective));
    result.l2_primal_residual              = cv.l2_primal_residual;
    result.l2_dual_residual                = cv.l2_dual_residual;
    result.l2_relative_primal_residual     = cv.l2_primal_residual / (1.0 + l2_norm_b);
    result.l2_relative_dual_residual       = cv.l2_dual_residual / (1.0 + l2_norm_c);
    result.step_size                       = step_size;
    result.primal_weight                   = primal_weight;
    result.n_restarts                      = n_restarts;
    result.n_major                         = n_major;
    result.solution_is_average             = is_avg ? 1 : 0;
  }

  // save_best_primal_so_far (pdlp.cu:333-463): quality = primal feasible first, then objective; else least l2 residual
  struct quality_t {
    bool feasible    = false;
    double residual  = INFINITY;
    double objective = INFINITY;  // -inf when maximising (pdlp.cu ctor)
  };
  quality_t best_quality;
  bool have_best = false;
  stats_t best_result{};
  std::vector<double> best_x, best_y, best_rc;
  bool first_is_better(const quality_t& a, const quality_t& b) const  // get_best_quality(current = a, other = b) == a
  {
    if (a.feasible && !b.feasible) return true;
    if (!a.feasible && b.feasible) return false;
    if (a.feasible && b.feasible) {
      const bool lower = a.objective < b.objective;
      return (!maximize && lower) || (maximize && !lower);
    }
    return a.residual < b.residual;
  }
  void record_best_primal_so_far()
  {
    if (!have_best && best_quality.objective == INFINITY && maximize) best_quality.objective = -INFINITY;
    const quality_t qc{conv_cur.status == 7, conv_cur.l2_primal_residual, conv_cur.primal_objective};
    const quality_t qa{conv_avg.status == 7, conv_avg.l2_primal_residual, conv_avg.primal_objective};
    const bool cur_wins      = first_is_better(qc, qa);
    const quality_t& cand    = cur_wins ? qc : qa;
    if (!first_is_better(cand, best_quality)) return;
    best_quality = cand;
    // fill_return_problem_solution at this moment (status is overwritten when a limit returns it)
    const stats_t keep_r = result;
    const auto keep_x = sol_x, keep_y = sol_y, keep_rc = sol_rc;
    if (cur_wins) fill_solution(x, y, conv_cur, 5, false);
    else fill_solution(x_avg, y_avg, conv_avg, 5, true);
    best_result = result; best_x = sol_x; best_y = sol_y; best_rc = sol_rc;
    result = keep_r; sol_x = keep_x; sol_y = keep_y; sol_rc = keep_rc;
    have_best = true;
  }
  bool return_best(int status)
  {
    if (!(st.save_best_primal_so_far && have_best)) return false;
    result = best_result;
    result.termination_status = status;
    sol_x = best_x; sol_y = best_y; sol_rc = best_rc;
    return true;
  }

  // pdlp.cu:265-331 (time limit is evaluated by the caller's clock)
  bool check_limits(double elapsed)
  {
    if (elapsed >= st.time_limit) {
      if (!return_best(5)) fill_solution(x, y, conv_cur, 5, false);
      return true;
    }
    if (k_internal >= st.iteration_limit) {
      if (!return_best(4)) fill_solution(x, y, conv_cur, 4, false);
      return true;
    }
    return false;
  }

  // pdlp.cu:538-802 (first_primal_feasible / save_best_primal_so_far / infeasibility detection off)
  bool check_termination(double elapsed)
  {
    compute_convergence_information(x, y, conv_cur);
    compute_convergence_information(x_avg, y_avg, conv_avg);
    if (k_total <= 1) return check_limits(elapsed);  // :580-583
    const bool cur_opt = conv_cur.status == 1, avg_opt = conv_avg.status == 1;
    if (cur_opt && avg_opt) {  // :636-682
      if (kkt_score(conv_cur) < kkt_score(conv_avg)) fill_solution(x, y, conv_cur, 1, false);
      else fill_solution(x_avg, y_avg, conv_avg, 1, true);
      return true;
    }
    if (avg_opt) { fill_solution(x_avg, y_avg, conv_avg, 1, true); return true; }   // :685-700
    if (cur_opt) { fill_solution(x, y, conv_cur, 1, false); return true; }          // :701-716
    // pdlp.cu:716-770: infeasibility.  strict: any of the two iterates suffices; else both must agree
    {
      const int sc = conv_cur.status, sa = conv_avg.status;
      const bool ic = sc == 2 || sc == 3, ia = sa == 2 || sa == 3;
      if (st.strict_infeasibility) {
        if (ic) { fill_solution(x, y, conv_cur, sc, false); return true; }
        if (ia) { fill_solution(x_avg, y_avg, conv_avg, sa, true); return true; }
      } else if (ic && sc == sa) {
        fill_solution(x, y, conv_cur, sc, false);
        return true;
      }
    }
    if (valid_step_size == -1) {  // :780-789: error solution (empty vectors)
      result                    = stats_t{};
      result.termination_status = 6;
      result.number_of_steps_taken           = k_internal;
      result.total_number_of_attempted_steps = k_pdhg_host;
      sol_x.assign(n, 0.0); sol_y.assign(m, 0.0); sol_rc.assign(n, 0.0);
      return true;
    }
    if (st.save_best_primal_so_far) record_best_primal_so_far();  // :790-796
    return check_limits(elapsed);
  }

  bool should_do_artificial_restart(int total_iterations) const  // pdlp_restart_strategy.cu:940-961
  {
    return its_since_restart >= hp.artificial_restart_threshold * total_iterations;
  }
  bool kkt_decay(double cand) const  // :408-429
  {
    if (cand < hp.sufficient_reduction_for_restart * last_restart_kkt) return true;
    if (cand < hp.necessary_reduction_for_restart * last_restart_kkt && cand > last_candidate_kkt) return true;
    return false;
  }

  // pdlp_restart_strategy.cu:468-641; x/y and x_avg/y_avg are in the space pdlp.cu:1144-1149 left them in
  // (scaled when rescale_for_restart, otherwise still unscaled — the reference's behaviour).
  void run_kkt_restart(trace_row_t& tr)
  {
    const double cur = kkt_score(conv_cur);
    tr.kkt_current   = cur;
    if (its_since_restart == 0) {  // :505-513
      last_candidate_kkt = cur;
      last_restart_kkt   = cur;
      return;
    }
    const double avg = kkt_score(conv_avg);
    tr.kkt_average   = avg;
    bool to_avg;
    double cand;
    if (cur < avg) { to_avg = false; cand = cur; } else { to_avg = true; cand = avg; }
    if (should_do_artificial_restart(k_total) || kkt_decay(cand)) {
      const bool use_avg               = to_avg && !hp.never_restart_to_average;
      const std::vector<double>& candx = use_avg ? x_avg : x;
      const std::vector<double>& candy = use_avg ? y_avg : y;
      // :753-801, 1681-1714: plain L2 distance to the last restart point
      const double dprim = reduce_sum(n, [&](int j) { const double d = x_lr[j] - candx[j]; return d * d; });
      const double ddual = reduce_sum(m, [&](int i) { const double d = y_lr[i] - candy[i]; return d * d; });
      if (use_avg) {  // :587-599
        x = x_avg;
        y = y_avg;
        last_restart_was_average = true;
      } else {
        last_restart_was_average = false;
      }
      auto new_primal_weight = [&]() {  // :685-732
        const double pd = std::sqrt(dprim), dd = std::sqrt(ddual);
        const double guard = 1.0e-10;
        if (pd < guard || pd >= 1.0 / guard || dd < guard || dd >= 1.0 / guard) return;
        const double est = dd / pd;
        const double lw  = hp.primal_weight_update_smoothing * std::log(est) +
                          (1.0 - hp.primal_weight_update_smoothing) * std::log(primal_weight);
        primal_weight = std::exp(lw);
        tau           = step_size / primal_weight;
        sigma         = step_size * primal_weight;
      };
      // :601-611: the two orders differ only in which primal weight feeds `distance_traveled`, a quantity
      // the KKT scheme never reads, so both collapse to: store the restart point, update the weight.
      x_lr = x;  // x already holds the candidate (either untouched current, or the average copied above)
      y_lr = y;
      new_primal_weight();
      std::fill(sum_x.begin(), sum_x.end(), 0.0);  // weighted_average_solution.cu:51-60
      std::fill(sum_y.begin(), sum_y.end(), 0.0);
      sum_w             = 0.0;
      its_since_restart = 0;
      last_restart_kkt  = cand;
      n_restarts += 1;
      tr.restarted  = 1;
      tr.to_average = use_avg ? 1 : 0;
    }
    last_candidate_kkt = cand;
  }

  // ------------------------------------------------ trust-region restart (Methodical1)
  // localized_duality_gap_container_t: a point, its distances from the last restart and the bounds
  // on the optimal objective inside the trust region around it.
  struct local_gap_t {
    std::vector<double> px, py;
    double primal_distance = 0, dual_distance = 0, distance = 0;  // squared L2 moves / weighted radius
    double lower_bound = 0, upper_bound = 0, normalized_gap = 0;
  };
  // pdlp_restart_strategy.cu:804-817
  double weighted_distance(const local_gap_t& g) const
  {
    return std::sqrt(g.primal_distance * hp.primal_distance_smoothing * primal_weight +
                     g.dual_distance * (hp.dual_distance_smoothing / primal_weight));
  }
  // :1681-1714 (plain L2 of the difference to the last restart point)
  void distance_from_last_restart(local_gap_t& g) const
  {
    g.primal_distance = reduce_sum(n, [&](int j) { const double d = x_lr[j] - g.px[j]; return d * d; });
    g.dual_distance   = reduce_sum(m, [&](int i) { const double d = y_lr[i] - g.py[i]; return d * d; });
    g.distance        = weighted_distance(g);
  }
  // bound_optimal_objective (:1034-1051): gradients (:1717-1825), Lagrangian (:1828-1900), trust-region
  // solve (:1391-1678) on the SCALED problem; radius = g.distance.
  void bound_optimal_objective(local_gap_t& g)
  {
    const int N = n + m;
    std::vector<double> gp(n), gd(m), sub(m), aty(n);
    spmv(ATs, g.py.data(), aty.data());
    for (int j = 0; j < n; ++j) gp[j] = cs[j] - aty[j];  // c - A^T y
    spmv(As, g.px.data(), gd.data());                   // primal product A x
    for (int i = 0; i < m; ++i) {                       // compute_subgradient_kernel :1746-1783
      const double lo = lcs[i], up = ucs[i], prod = gd[i], yi = g.py[i];
      double sc;
      if (yi < 0.0) sc = up;
      else if (yi > 0.0) sc = lo;
      else if (!std::isfinite(up) && !std::isfinite(lo)) sc = 0.0;
      else if (!std::isfinite(up) && std::isfinite(lo)) sc = lo;
      else if (std::isfinite(up) && !std::isfinite(lo)) sc = up;
      else sc = prod < lo ? lo : (prod > up ? up : prod);
      sub[i] = sc;
      gd[i]  = sc - prod;  // dual gradient = subgradient - A x
    }
    const double lagrangian = reduce_sum(n, [&](int j) { return g.px[j] * cs[j]; }) -
                              reduce_sum(n, [&](int j) { return g.px[j] * aty[j]; }) +
                              reduce_sum(m, [&](int i) { return g.py[i] * sub[i]; });
    // ---- solve_bound_constrained_trust_region ----
    std::vector<double> center(N), obj(N), lo(N), up(N), w(N), dir(N, 0.0), thr(N, 0.0);
    for (int j = 0; j < n; ++j) { center[j] = g.px[j]; obj[j] = gp[j]; lo[j] = ls[j]; up[j] = us[j]; w[j] = 1.0 / tau; }
    for (int i = 0; i < m; ++i) {
      center[n + i] = g.py[i];
      obj[n + i]    = -gd[i];
      lo[n + i]     = std::isfinite(ucs[i]) ? -INFINITY : 0.0;  // utils.cuh:240-255
      up[n + i]     = std::isfinite(lcs[i]) ? INFINITY : 0.0;
      w[n + i]      = 1.0 / sigma;
    }
    const double obj_norm = std::sqrt(reduce_sum(N, [&](int k) { return obj[k] * obj[k]; }));
    std::vector<double> tr(center);
    if (!(g.distance == 0.0 || obj_norm == 0.0)) {
      for (int k = 0; k < N; ++k) {  // compute_direction_and_threshold, utils.cuh:291-322
        if (center[k] >= up[k] && obj[k] <= 0.0) continue;
        if (center[k] <= lo[k] && obj[k] >= 0.0) continue;
        if (obj[k] == 0.0) { thr[k] = INFINITY; continue; }
        dir[k] = -obj[k] / w[k];
        if (dir[k] > 0.0) thr[k] = (up[k] - center[k]) / dir[k];
        else if (dir[k] < 0.0) thr[k] = (lo[k] - center[k]) / dir[k];
      }
      double high_r2 = 0.0, low_r2 = 0.0;
      for (int k = 0; k < N; ++k)
        if (std::isinf(thr[k])) high_r2 += dir[k] * dir[k] * w[k];  // weighted_l2_if_infinite
      std::vector<int> order(N);
      for (int k = 0; k < N; ++k) order[k] = k;
      std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return thr[a] < thr[b]; });
      int range_low = 0, range_high = N;
      while (range_low < N && thr[order[range_low]] == -INFINITY) ++range_low;
      for (int k = 0; k < N; ++k)
        if (thr[order[k]] == INFINITY) { range_high = k; break; }
      const double target = g.distance;
      std::vector<double> test(N, 0.0);
      auto sthr = [&](int k) { return thr[order[k]]; };
      while (range_low != range_high) {  // solve_bound_constrained_trust_region_kernel :1291-1358
        const int size = range_high - range_low;
        const double t = (size & 1) == 0 ? 0.5 * (sthr(range_low + size / 2 - 1) + sthr(range_low + size / 2))
                                         : sthr(range_low + size / 2);
        double test_r2 = 0.0;
        for (int k = range_low; k < range_high; ++k) {
          const int q = order[k];
          test[k]     = std::min(std::max(center[q] + t * dir[q], lo[q]), up[q]);
          const double d = test[k] - center[q];
          test_r2 += d * d * w[q];
        }
        const bool too_high = low_r2 + test_r2 + (t * t) * high_r2 >= target * target;
        if (too_high) {
          int new_high = range_high;
          for (int k = range_low; k < range_high; ++k)
            if (sthr(k) >= t) { new_high = k; break; }
          for (int k = new_high; k < range_high; ++k) high_r2 += dir[order[k]] * dir[order[k]] * w[order[k]];
          range_high = new_high;
        } else {
          int new_low = range_low;
          for (int k = range_high - 1; k >= range_low; --k)
            if (sthr(k) <= t) { new_low = k + 1; break; }
          for (int k = range_low; k < new_low; ++k) {
            const double d = test[k] - center[order[k]];
            low_r2 += d * d * w[order[k]];
          }
          range_low = new_low;
        }
      }
      double target_threshold;  // target_threshold_determination_kernel :1081-1100
      if (high_r2 <= 0.0) target_threshold = *std::max_element(thr.begin(), thr.end());
      else target_threshold = std::sqrt((target * target - low_r2) / high_r2);
      for (int k = 0; k < N; ++k) {
        // a component that does not move keeps its value (inf * 0 would be NaN in the reference's a + t * b)
        const double moved = dir[k] == 0.0 ? center[k] : center[k] + target_threshold * dir[k];
        tr[k]              = std::min(std::max(moved, lo[k]), up[k]);
      }
    }
    // compute_bound :1053-1078
    g.lower_bound = lagrangian + reduce_sum(n, [&](int j) { return (tr[j] - g.px[j]) * gp[j]; });
    g.upper_bound = lagrangian + reduce_sum(m, [&](int i) { return (tr[n + i] - g.py[i]) * gd[i]; });
  }

  // pdlp_restart_strategy.cu:278-364
  void run_trust_region_restart(trace_row_t& tr)
  {
    if (its_since_restart == 0) return;
    bool restart = should_do_artificial_restart(k_total);
    // compute_localized_duality_gaps :983-1031
    local_gap_t avg, cur;
    avg.px = x_avg; avg.py = y_avg;
    cur.px = x;     cur.py = y;
    distance_from_last_restart(avg);
    distance_from_last_restart(cur);
    bound_optimal_objective(avg);
    bound_optimal_objective(cur);
    avg.normalized_gap = (avg.upper_bound - avg.lower_bound) / avg.distance;
    cur.normalized_gap = (cur.upper_bound - cur.lower_bound) / cur.distance;
    // pick_restart_candidate :842-873
    const bool to_avg = cur.normalized_gap / cur.distance >= avg.normalized_gap / avg.distance;
    local_gap_t& cand = to_avg ? avg : cur;
    if (!restart) {  // should_do_adaptive_restart_normalized_duality_gap :903-937
      local_gap_t last;
      last.px = x_lr; last.py = y_lr;
      last.primal_distance = cand.primal_distance;  // only the radius is taken from the candidate (:921-926)
      last.dual_distance   = cand.dual_distance;
      last.distance        = weighted_distance(cand);
      bound_optimal_objective(last);
      last.normalized_gap = (last.upper_bound - last.lower_bound) / last.distance;
      const double ratio  = cand.normalized_gap / last.normalized_gap;
      if (ratio < hp.necessary_reduction_for_restart &&
          (ratio < hp.sufficient_reduction_for_restart || ratio > gap_reduction_ratio_last_trial))
        restart = true;
      gap_reduction_ratio_last_trial = ratio;
    }
    if (!restart) return;
    const bool use_avg = to_avg && !hp.never_restart_to_average;
    if (use_avg) { x = cand.px; y = cand.py; }
    last_restart_was_average = use_avg;
    auto new_primal_weight = [&]() {  // :685-732
      const double pd = std::sqrt(cand.primal_distance), dd = std::sqrt(cand.dual_distance);
      const double guard = 1.0e-10;
      if (pd < guard || pd >= 1.0 / guard || dd < guard || dd >= 1.0 / guard) return;
      const double lw = hp.primal_weight_update_smoothing * std::log(dd / pd) +
                        (1.0 - hp.primal_weight_update_smoothing) * std::log(primal_weight);
      primal_weight = std::exp(lw);
      tau           = step_size / primal_weight;
      sigma         = step_size * primal_weight;
    };
    // :341-349 — both orders store the candidate as the new restart point; the weight only feeds the stored radius,
    // which the next call recomputes
    x_lr = cand.px;
    y_lr = cand.py;
    new_primal_weight();
    std::fill(sum_x.begin(), sum_x.end(), 0.0);
    std::fill(sum_y.begin(), sum_y.end(), 0.0);
    sum_w             = 0.0;
    its_since_restart = 0;
    n_restarts += 1;
    tr.restarted  = 1;
    tr.to_average = use_avg ? 1 : 0;
  }

  // -------------------------------------------------------- outer loop
  // pdlp.cu:1081-1185.  Runs until termination or until `max_accepted_steps` more PDLP iterations
  // have been taken (used by the step-by-step parity tests).  Returns true when a solution was filled.
  bool run(int max_accepted_steps)
  {
    if (!initialised) initialise();
    const auto t0 = std::chrono::steady_clock::now();
    int budget    = max_accepted_steps;
    while (true) {
      const bool is_major = ((k_total % hp.major_iteration == 0) && (k_total > 0)) || (k_total <= hp.min_iteration_restart);
      const bool error    = (valid_step_size == -1);
      bool artificial     = false;
      if (hp.artificial_restart_in_main_loop) artificial = should_do_artificial_restart(k_total);
      if (is_major || artificial || error) {
        n_major += 1;
        trace_row_t tr{};
        tr.k = k_total;
        const bool no_rescale_average = k_internal == 0 && warm_given;  // :1100-1101
        if (no_rescale_average) {
          // the averages handed in with the warm start are already unscaled
        } else if (k_internal <= 1) {  // :1110-1118
          x_avg = x;
          y_avg = y;
        } else {  // weighted_average_solution.cu:114-142
          if (its_since_restart == 0) {
            std::fill(x_avg.begin(), x_avg.end(), 0.0);
            std::fill(y_avg.begin(), y_avg.end(), 0.0);
          } else {
            for (int j = 0; j < n; ++j) x_avg[j] = sum_x[j] / sum_w;
            for (int i = 0; i < m; ++i) y_avg[i] = sum_y[i] / sum_w;
          }
        }
        if (!no_rescale_average) unscale_solutions(x_avg, y_avg);  // :1131-1136
        unscale_solutions(x, y);
        const double elapsed = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        const bool done      = check_termination(elapsed);
        tr.primal_objective = conv_cur.primal_objective; tr.dual_objective = conv_cur.dual_objective;
        tr.gap = conv_cur.gap; tr.l2_primal_residual = conv_cur.l2_primal_residual;
        tr.l2_dual_residual = conv_cur.l2_dual_residual; tr.step_size = step_size; tr.primal_weight = primal_weight;
        if (done) {
          trace.push_back(tr);
          result.solve_seconds = elapsed;
          return true;
        }
        if (hp.rescale_for_restart) {  // :1144-1149
          scale_solutions(x_avg, y_avg);
          scale_solutions(x, y);
        }
        if (hp.restart_strategy == 1) run_kkt_restart(tr);
        else if (hp.restart_strategy == 2) run_trust_region_restart(tr);
        if (!hp.rescale_for_restart) scale_solutions(x, y);  // :1168-1175
        trace.push_back(tr);
      }
      if (budget == 0) return false;
      take_step();
      ++k_total;
      ++k_internal;
      if (budget > 0) --budget;
    }
  }
};

}  // namespace

// ------------------------------------------------------------------ C ABI (ctypes)
extern "C" {

void* pdlp_oracle_create(int m, int n, const int* off, const int* idx, const double* val, const double* c,
                         const double* l, const double* u, const double* lc, const double* uc, int maximize,
                         double objective_offset, const hyper_t* hp, const settings_t* st)
{
#ifdef _OPENMP
  if (st->num_threads > 0) omp_set_num_threads(st->num_threads);  // results do not depend on it (chunked reductions)
#endif
  return new oracle_t(m, n, off, idx, val, c, l, u, lc, uc, maximize, objective_offset, *hp, *st);
}
void pdlp_oracle_destroy(void* h) { delete static_cast<oracle_t*>(h); }
void pdlp_oracle_initialise(void* h) { static_cast<oracle_t*>(h)->initialise(); }
// returns 1 if terminated (solution filled), 0 if the step budget ran out first. max_steps < 0: run to the end.
int pdlp_oracle_run(void* h, int max_steps) { return static_cast<oracle_t*>(h)->run(max_steps) ? 1 : 0; }
void pdlp_oracle_stats(void* h, stats_t* out) { *out = static_cast<oracle_t*>(h)->result; }

// One PDHG attempt from caller-supplied scaled state (kernel-level parity tests): computes
// x', xbar, y', AtY', the three reductions; no accept/reject, no state mutation besides the buffers.
void pdlp_oracle_single_attempt(void* h, const double* x, const double* y, const double* aty, double tau, double sigma,
                                double* xn, double* xbar, double* yn, double* atyn, double* red3)
{
  auto* o = static_cast<oracle_t*>(h);
  oracle_t& s = *o;
  std::vector<double> sx = s.x, sy = s.y, saty = s.AtY;
  const double stau = s.tau, ssig = s.sigma;
  s.x.assign(x, x + s.n); s.y.assign(y, y + s.m); s.AtY.assign(aty, aty + s.n);
  s.tau = tau; s.sigma = sigma;
  s.primal_projection();
  s.dual_projection();
  s.interaction_and_movement();
  std::memcpy(xn, s.xn.data(), sizeof(double) * s.n);
  std::memcpy(xbar, s.xbar.data(), sizeof(double) * s.n);
  std::memcpy(yn, s.yn.data(), sizeof(double) * s.m);
  std::memcpy(atyn, s.AtYn.data(), sizeof(double) * s.n);
  red3[0] = s.interaction; red3[1] = s.norm_dx2; red3[2] = s.norm_dy2;
  s.x = sx; s.y = sy; s.AtY = saty; s.tau = stau; s.sigma = ssig;
}

// Evaluate the termination quantities for an UNSCALED iterate: out[0..7] = l2_primal_residual, l2_dual_residual,
// primal_objective, dual_objective, gap, abs_objective, status, kkt(primal_weight as currently held)
void pdlp_oracle_convergence(void* h, const double* px, const double* py, double* out8, double* reduced_cost)
{
  auto* o = static_cast<oracle_t*>(h);
  if (o->tmp_m.size() != (size_t)o->m) { o->tmp_m.assign(o->m, 0.0); o->tmp_n.assign(o->n, 0.0); }
  convergence_t cv;
  cv.reduced_cost.assign(o->n, 0.0);
  std::vector<double> vx(px, px + o->n), vy(py, py + o->m);
  o->compute_convergence_information(vx, vy, cv);
  out8[0] = cv.l2_primal_residual; out8[1] = cv.l2_dual_residual; out8[2] = cv.primal_objective;
  out8[3] = cv.dual_objective; out8[4] = cv.gap; out8[5] = cv.abs_objective; out8[6] = cv.status;
  out8[7] = o->primal_weight > 0 ? o->kkt_score(cv) : 0.0;
  o->last_linf[0] = cv.linf_relative_primal_residual;
  o->last_linf[1] = cv.linf_relative_dual_residual;
  if (reduced_cost) std::memcpy(reduced_cost, cv.reduced_cost.data(), sizeof(double) * o->n);
}

// Warm start (pdlp.cu:469-489 produce / :131-181 consume).  get: valid after a run() that returned 1 (x, y and the
// averages are unscaled at that point, everything else scaled — exactly what the reference hands out).
void pdlp_oracle_get_warm_start(void* h, double* const* vectors9, double* scalars8)
{
  auto* o = static_cast<oracle_t*>(h);
  const std::vector<double>* v[9] = {&o->x, &o->y, &o->x_avg, &o->y_avg, &o->AtY, &o->sum_x, &o->sum_y, &o->x_lr, &o->y_lr};
  for (int q = 0; q < 9; ++q) std::memcpy(vectors9[q], v[q]->data(), sizeof(double) * v[q]->size());
  scalars8[0] = o->primal_weight; scalars8[1] = o->step_size; scalars8[2] = o->k_total; scalars8[3] = o->k_pdhg_host;
  scalars8[4] = o->last_candidate_kkt; scalars8[5] = o->last_restart_kkt; scalars8[6] = o->sum_w;
  scalars8[7] = o->its_since_restart;
}
// set: before initialise() / the first run()
void pdlp_oracle_set_warm_start(void* h, const double* const* vectors9, const double* scalars8)
{
  auto* o = static_cast<oracle_t*>(h);
  const bool primal[9] = {true, false, true, false, true, true, false, true, false};
  for (int q = 0; q < 9; ++q) o->warm_v[q].assign(vectors9[q], vectors9[q] + (primal[q] ? o->n : o->m));
  for (int q = 0; q < 8; ++q) o->warm_s[q] = scalars8[q];
  o->warm_given = true;
}

// Trust-region bounds at a caller-supplied point of the SCALED space (white-box test of the device reformulation):
// out4 = {lower bound, upper bound, weighted distance used as radius, 0}; radius < 0 -> distance to the last restart.
void pdlp_oracle_trust_region_bounds(void* h, const double* px, const double* py, double radius, double* out4)
{
  auto* o = static_cast<oracle_t*>(h);
  oracle_t::local_gap_t g;
  g.px.assign(px, px + o->n);
  g.py.assign(py, py + o->m);
  o->distance_from_last_restart(g);
  if (radius >= 0.0) g.distance = radius;
  o->bound_optimal_objective(g);
  out4[0] = g.lower_bound; out4[1] = g.upper_bound; out4[2] = g.distance; out4[3] = 0.0;
}

// Named vectors / scalars for white-box comparisons.
int pdlp_oracle_get_vector(void* h, const char* name, double* out)
{
  auto* o = static_cast<oracle_t*>(h);
  const std::string s(name);
  const std::vector<double>* v = nullptr;
  if (s == "x") v = &o->x; else if (s == "y") v = &o->y; else if (s == "aty") v = &o->AtY;
  else if (s == "sum_x") v = &o->sum_x; else if (s == "sum_y") v = &o->sum_y;
  else if (s == "x_avg") v = &o->x_avg; else if (s == "y_avg") v = &o->y_avg;
  else if (s == "row_scaling") v = &o->Dr; else if (s == "col_scaling") v = &o->Dc;
  else if (s == "scaled_values") v = &o->As.val; else if (s == "scaled_values_t") v = &o->ATs.val;
  else if (s == "scaled_c") v = &o->cs; else if (s == "scaled_l") v = &o->ls; else if (s == "scaled_u") v = &o->us;
  else if (s == "scaled_lc") v = &o->lcs; else if (s == "scaled_uc") v = &o->ucs;
  else if (s == "solution_x") v = &o->sol_x; else if (s == "solution_y") v = &o->sol_y;
  else if (s == "solution_rc") v = &o->sol_rc;
  else if (s == "x_last_restart") v = &o->x_lr; else if (s == "y_last_restart") v = &o->y_lr;
  if (!v) return -1;
  if (out) std::memcpy(out, v->data(), sizeof(double) * v->size());
  return (int)v->size();
}
double pdlp_oracle_get_scalar(void* h, const char* name)
{
  auto* o = static_cast<oracle_t*>(h);
  const std::string s(name);
  if (s == "step_size") return o->step_size;
  if (s == "primal_weight") return o->primal_weight;
  if (s == "tau") return o->tau;
  if (s == "sigma") return o->sigma;
  if (s == "sum_w") return o->sum_w;
  if (s == "k_total") return o->k_total;
  if (s == "k_pdhg") return o->k_pdhg_host;
  if (s == "its_since_restart") return o->its_since_restart;
  if (s == "interaction") return o->interaction;
  if (s == "norm_dx2") return o->norm_dx2;
  if (s == "norm_dy2") return o->norm_dy2;
  if (s == "l2_norm_b") return o->l2_norm_b;
  if (s == "l2_norm_c") return o->l2_norm_c;
  if (s == "last_restart_kkt") return o->last_restart_kkt;
  if (s == "last_candidate_kkt") return o->last_candidate_kkt;
  if (s == "n_restarts") return o->n_restarts;
  if (s == "linf_relative_primal_residual") return o->last_linf[0];
  if (s == "linf_relative_dual_residual") return o->last_linf[1];
  return std::numeric_limits<double>::quiet_NaN();
}
// Major-iteration trace: 12 doubles per row (k, restarted, to_average, p, d, gap, rp, rd, step, weight, kkt_cur, kkt_avg)
int pdlp_oracle_trace(void* h, double* out, int max_rows)
{
  auto* o   = static_cast<oracle_t*>(h);
  const int r = (int)o->trace.size();
  if (out) {
    for (int i = 0; i < std::min(r, max_rows); ++i) {
      const auto& t = o->trace[i];
      double* p     = out + 12 * i;
      p[0] = t.k; p[1] = t.restarted; p[2] = t.to_average; p[3] = t.primal_objective; p[4] = t.dual_objective;
      p[5] = t.gap; p[6] = t.l2_primal_residual; p[7] = t.l2_dual_residual; p[8] = t.step_size;
      p[9] = t.primal_weight; p[10] = t.kkt_current; p[11] = t.kkt_average;
    }
  }
  return r;
}

}  // extern "C"

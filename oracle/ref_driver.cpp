// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// Thin C-ABI shim around two pieces of the *unmodified* reference that compile
// with plain g++ (no RAFT/RMM/cmake):
//   * cpp/libmps_parser            (MPS -> host CSR data model)
//   * cpp/src/dual_simplex         (the reference's CPU LP path)
// The reference sources are compiled where they lie under /root/reference by
// oracle/Makefile; only this shim lives in the repo.  Output goes to
// oracle/_ref/libcuopt_ref_cpu.so (git-ignored, travels to the GPU box).
//
// Used by: tests/ (parser parity, known-answer objectives), the golden-vector
// generator scripts/gen_golden.py, and bench.py's cpu_baseline / --impl reference.
//
// The row-type/range translation below follows the reference's own hand-off to
// its simplex code: cpp/src/linear_programming/translate.hpp:30-100.

#include <mps_parser/parser.hpp>

#include <dual_simplex/solve.hpp>
#include <dual_simplex/sparse_matrix.hpp>
#include <dual_simplex/tic_toc.hpp>
#include <dual_simplex/user_problem.hpp>

#include <cmath>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

using model_t = cuopt::mps_parser::mps_data_model_t<int, double>;
namespace ds  = cuopt::linear_programming::dual_simplex;

extern "C" {

// ---- MPS parsing through the reference parser -----------------------------
void* ref_mps_parse(const char* path, int fixed_format, char* err, int errlen)
{
  try {
    auto* m = new model_t(cuopt::mps_parser::parse_mps<int, double>(std::string(path), fixed_format != 0));
    return m;
  } catch (const std::exception& e) {
    if (err && errlen > 0) { std::snprintf(err, errlen, "%s", e.what()); }
    return nullptr;
  }
}

void ref_mps_free(void* h) { delete static_cast<model_t*>(h); }

void ref_mps_dims(void* h, int* m, int* n, int* nnz, int* maximize, double* offset, double* scale)
{
  auto* d   = static_cast<model_t*>(h);
  *m        = d->get_n_constraints();
  *n        = d->get_n_variables();
  *nnz      = d->get_nnz();
  *maximize = d->get_sense() ? 1 : 0;
  *offset   = d->get_objective_offset();
  *scale    = d->get_objective_scaling_factor();
}

// Every pointer may be null (skipped).  Sizes are the caller's responsibility (from ref_mps_dims).
void ref_mps_arrays(void* h,
                    int* offsets,
                    int* indices,
                    double* values,
                    double* rhs,
                    double* c,
                    double* var_lb,
                    double* var_ub,
                    double* con_lb,
                    double* con_ub,
                    char* row_types,
                    char* var_types)
{
  auto* d = static_cast<model_t*>(h);
  auto cp = [](auto* dst, const auto& src) {
    if (dst && !src.empty()) std::memcpy(dst, src.data(), src.size() * sizeof(src[0]));
  };
  cp(offsets, d->get_constraint_matrix_offsets());
  cp(indices, d->get_constraint_matrix_indices());
  cp(values, d->get_constraint_matrix_values());
  cp(rhs, d->get_constraint_bounds());
  cp(c, d->get_objective_coefficients());
  cp(var_lb, d->get_variable_lower_bounds());
  cp(var_ub, d->get_variable_upper_bounds());
  cp(con_lb, d->get_constraint_lower_bounds());
  cp(con_ub, d->get_constraint_upper_bounds());
  cp(row_types, d->get_row_types());
  cp(var_types, d->get_variable_types());
}

// Names joined with '\n' (var names when which==0, row names when which==1,
// "problem\nobjective" when which==2).  Returns required length incl. NUL.
int ref_mps_names(void* h, int which, char* out, int outlen)
{
  auto* d = static_cast<model_t*>(h);
  std::string s;
  if (which == 2) {
    s = d->get_problem_name() + "\n" + d->get_objective_name();
  } else {
    const auto& v = which == 0 ? d->get_variable_names() : d->get_row_names();
    for (size_t i = 0; i < v.size(); ++i) {
      if (i) s += '\n';
      s += v[i];
    }
  }
  if (out && outlen > 0) std::snprintf(out, outlen, "%s", s.c_str());
  return (int)s.size() + 1;
}

// ---- the reference's CPU dual simplex on an in-memory ranged LP ------------
// min/max  scale * c'x + offset   s.t.  con_lb <= A x <= con_ub,  var_lb <= x <= var_ub
// `obj_scale` is +1 for minimise, -1 for maximise with `c` ALREADY negated by
// the caller when maximising (reference convention, problem_helpers.cuh:128-142).
// status: dual_simplex::lp_status_t as int.  Returns 0 on success, -1 on exception.
int ref_dual_simplex(int m,
                     int n,
                     const int* row_offsets,
                     const int* col_indices,
                     const double* values,
                     const double* con_lb,
                     const double* con_ub,
                     const double* c,
                     const double* var_lb,
                     const double* var_ub,
                     double obj_scale,
                     double obj_offset,
                     double time_limit,
                     int iteration_limit,
                     int log,
                     double* x,
                     double* y,
                     double* z,
                     double* user_objective,
                     int* iterations,
                     int* status,
                     double* seconds)
{
  try {
    const double inf = std::numeric_limits<double>::infinity();
    ds::user_problem_t<int, double> up;
    up.num_rows = m;
    up.num_cols = n;
    up.objective.assign(c, c + n);
    ds::csr_matrix_t<int, double> csr;
    const int nz  = row_offsets[m];
    csr.m         = m;
    csr.n         = n;
    csr.nz_max    = nz;
    csr.x.assign(values, values + nz);
    csr.j.assign(col_indices, col_indices + nz);
    csr.row_start.assign(row_offsets, row_offsets + m + 1);
    csr.to_compressed_col(up.A);
    up.rhs.resize(m);
    up.row_sense.resize(m);
    for (int i = 0; i < m; ++i) {
      const double lo = con_lb[i], hi = con_ub[i];
      if (lo == hi) {
        up.row_sense[i] = 'E';
        up.rhs[i]       = lo;
      } else if (hi == inf) {
        up.row_sense[i] = 'G';
        up.rhs[i]       = lo;
      } else if (lo == -inf) {
        up.row_sense[i] = 'L';
        up.rhs[i]       = hi;
      } else {
        up.row_sense[i] = 'E';
        up.rhs[i]       = lo;
        up.range_rows.push_back(i);
        up.range_value.push_back(hi - lo);
      }
    }
    up.num_range_rows = (int)up.range_rows.size();
    up.lower.assign(var_lb, var_lb + n);
    up.upper.assign(var_ub, var_ub + n);
    up.obj_constant = obj_offset;
    up.obj_scale    = obj_scale;
    up.var_types.assign(n, ds::variable_type_t::CONTINUOUS);

    ds::simplex_solver_settings_t<int, double> settings;
    settings.time_limit      = time_limit;
    settings.iteration_limit = iteration_limit;
    settings.log.log         = log != 0;
    ds::lp_solution_t<int, double> sol(m, n);
    const double t0 = ds::tic();
    auto st         = ds::solve_linear_program<int, double>(up, settings, sol);
    *seconds        = ds::toc(t0);
    *status         = static_cast<int>(st);
    *iterations     = sol.iterations;
    *user_objective = sol.user_objective;
    if (x) std::memcpy(x, sol.x.data(), sizeof(double) * n);
    if (y) std::memcpy(y, sol.y.data(), sizeof(double) * m);
    if (z) std::memcpy(z, sol.z.data(), sizeof(double) * n);
    return 0;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "ref_dual_simplex: %s\n", e.what());
    return -1;
  }
}

}  // extern "C"

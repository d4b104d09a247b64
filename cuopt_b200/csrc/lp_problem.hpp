// Host-side LP container behind the cuOptOptimizationProblem handle.
//
// Plays the role of the reference's optimization_problem_t
// (cpp/include/cuopt/linear_programming/optimization_problem.hpp) for the LP
// path only: it stores exactly what the caller handed to cuOptCreateProblem /
// cuOptCreateRangedProblem / cuOptReadProblem, on the HOST; the device layout
// (scaled CSR + transposed CSR, bound vectors) is built by pdlp_solver_t when
// cuOptSolve runs, so no CUDA call happens before a solve.
#pragma once

#include <cmath>
#include <limits>
#include <stdexcept>
#include <string>
#include <vector>
#include <atomic>
#include <algorithm>
#include <cstring>
#include <thread>

namespace cuopt_b200 {

// Mirrors cuopt::error_type_t / the CUOPT_* status codes (constants.h:107-113).
enum class error_type_t : int {
  Success         = 0,
  InvalidArgument = 1,
  MpsFileError    = 2,
  MpsParseError   = 3,
  ValidationError = 4,
  OutOfMemory     = 5,
  RuntimeError    = 6
};

struct lp_error : public std::runtime_error {
  lp_error(error_type_t t, const std::string& what) : std::runtime_error(what), type(t) {}
  error_type_t type;
};

// Host-side helpers for 10M-row problems: the C ABI copies every array it is given (cuopt_c.cpp:115-135) and validates
// the CSR; single-threaded that is ~0.4 s of a 1.4 GB problem, so both run on a few worker threads.
template <typename F>
inline void parallel_chunks(size_t n, F&& body, size_t min_chunk = size_t(1) << 20)
{
  const size_t hw   = std::max(1u, std::min(std::thread::hardware_concurrency(), 16u));
  const size_t n_th = std::max<size_t>(1, std::min(hw, n / min_chunk));
  if (n_th <= 1) {
    body(size_t(0), n);
    return;
  }
  std::vector<std::thread> pool;
  const size_t per = (n + n_th - 1) / n_th;
  for (size_t t = 0; t < n_th; ++t) {
    const size_t lo = t * per, hi = std::min(n, lo + per);
    if (lo >= hi) break;
    pool.emplace_back([&body, lo, hi]() { body(lo, hi); });
  }
  for (auto& th : pool) th.join();
}
// Vector whose resize() leaves trivially constructible elements uninitialised: a 640 MB array is then first touched by
// the worker threads that fill it, not zeroed by one thread first.
template <typename T>
struct noinit_allocator : std::allocator<T> {
  template <typename U>
  struct rebind {
    using other = noinit_allocator<U>;
  };
  template <typename U, typename... Args>
  void construct(U* p, Args&&... args)
  {
    if constexpr (sizeof...(Args) == 0) ::new (static_cast<void*>(p)) U;
    else ::new (static_cast<void*>(p)) U(std::forward<Args>(args)...);
  }
};
template <typename T>
using hvec = std::vector<T, noinit_allocator<T>>;

template <typename T>
inline void parallel_assign(hvec<T>& dst, const T* src, size_t n)
{
  dst.resize(n);
  if (n == 0) return;
  T* d = dst.data();
  parallel_chunks(n, [d, src](size_t lo, size_t hi) { std::memcpy(d + lo, src + lo, (hi - lo) * sizeof(T)); },
                  (size_t(4) << 20) / sizeof(T));
}

struct lp_problem_t {
  int n_constraints = 0;
  int n_variables   = 0;
  bool maximize     = false;
  double objective_offset         = 0.0;
  double objective_scaling_factor = 1.0;

  // CSR constraint matrix A (n_constraints x n_variables)
  hvec<int> A_offsets{0};
  hvec<int> A_indices;
  hvec<double> A_values;

  hvec<double> objective_coefficients;  // c
  hvec<double> variable_lower_bounds;   // may be empty -> 0
  hvec<double> variable_upper_bounds;   // may be empty -> +inf
  std::vector<char> variable_types;            // 'C' / 'I', may be empty -> all continuous

  // "sense" form (cuOptCreateProblem, MPS): row type 'E'/'G'/'L' + right-hand side
  std::vector<char> row_types;
  hvec<double> constraint_bounds;  // b
  // "ranged" form (cuOptCreateRangedProblem, MPS after RANGES): lc <= A x <= uc
  hvec<double> constraint_lower_bounds;
  hvec<double> constraint_upper_bounds;

  std::string problem_name, objective_name;
  std::vector<std::string> variable_names, row_names;

  int nnz() const { return (int)A_values.size(); }

  bool is_mip() const
  {
    for (char t : variable_types)
      if (t == 'I') return true;
    return false;
  }

  // Representation checks, same rules and same error category as
  // cpp/src/linear_programming/utilities/problem_checking.cu:30-250 (ValidationError).
  void check_representation() const
  {
    auto fail = [](const std::string& s) { throw lp_error(error_type_t::ValidationError, s); };
    const bool empty_problem = A_values.empty();
    if (A_offsets.empty()) fail("A_offsets must be set before calling the solver.");
    if (!empty_problem && objective_coefficients.empty()) fail("c must be set before calling the solver.");
    if (A_indices.size() != A_values.size()) fail("A_index and A_values must have same sizes.");
    if (A_offsets.front() != 0) fail("A_offsets first value should be 0.");
    for (size_t i = 1; i < A_offsets.size(); ++i)
      if (A_offsets[i] < A_offsets[i - 1]) fail("A_offsets values must in an increasing order.");
    if ((size_t)A_offsets.back() != A_values.size()) fail("A_offsets last value must equal the number of nonzeros.");
    {
      std::atomic<bool> bad{false};
      const int* idx = A_indices.data();
      const int nv   = n_variables;
      parallel_chunks(A_indices.size(), [&bad, idx, nv](size_t lo, size_t hi) {
        bool b = false;
        for (size_t p = lo; p < hi; ++p) b |= (idx[p] < 0) | (idx[p] >= nv);
        if (b) bad = true;
      });
      if (bad) fail("A_indices values must positive lower than the number of variables (c size).");
    }
    if (constraint_lower_bounds.empty() != constraint_upper_bounds.empty())
      fail("Constraints lower bounds must be set along with constraints upper bounds.");
    const bool have_ranged = !constraint_lower_bounds.empty();
    const bool have_sense  = !row_types.empty() && !constraint_bounds.empty();
    if (!empty_problem && !have_ranged && !have_sense)
      fail(
        "Either constraints lower/upper bounds or row types and constraints bounds needs to be set before calling "
        "the solver.");
    if (!row_types.empty()) {
      for (char t : row_types)
        if (t != 'E' && t != 'G' && t != 'L') fail("row_types values must equal to 'E', 'G' or 'L'.");
      if (row_types.size() != constraint_bounds.size())
        fail("Sizes for vectors related to the constraints are not the same (row types vs right hand side).");
      if (A_offsets.size() - 1 != constraint_bounds.size())
        fail("Sizes for vectors related to the constraints are not the same (right hand side vs matrix rows).");
    }
    if (have_ranged) {
      if (constraint_lower_bounds.size() != constraint_upper_bounds.size() ||
          constraint_lower_bounds.size() != A_offsets.size() - 1)
        fail("Sizes for vectors related to the constraints are not the same (constraint bounds vs matrix rows).");
    }
    if (!variable_lower_bounds.empty() && variable_lower_bounds.size() != objective_coefficients.size())
      fail("Sizes for vectors related to the variables are not the same (lower bounds vs objective).");
    if (!variable_upper_bounds.empty() && variable_upper_bounds.size() != objective_coefficients.size())
      fail("Sizes for vectors related to the variables are not the same (upper bounds vs objective).");
  }

  // Two-sided row bounds as PDLP uses them.  Rule of the reference's problem_t
  // construction (cpp/src/mip/problem/problem_helpers.cuh:34-87): explicit lower/upper
  // bounds win; otherwise E -> [b,b], G -> [b,+inf], L -> [-inf,b].
  void row_bounds(hvec<double>& lo, hvec<double>& hi) const
  {
    const double inf = std::numeric_limits<double>::infinity();
    if (!constraint_lower_bounds.empty()) {
      lo = constraint_lower_bounds;
      hi = constraint_upper_bounds;
      return;
    }
    lo.resize(row_types.size());
    hi.resize(row_types.size());
    for (size_t i = 0; i < row_types.size(); ++i) {
      const double b = constraint_bounds[i];
      switch (row_types[i]) {
        case 'E': lo[i] = b; hi[i] = b; break;
        case 'G': lo[i] = b; hi[i] = inf; break;
        case 'L': lo[i] = -inf; hi[i] = b; break;
        default: lo[i] = -inf; hi[i] = -inf; break;
      }
    }
  }

  // Default variable bounds [0, +inf) when not given (problem_helpers.cuh:89-116).
  void variable_bounds(hvec<double>& lo, hvec<double>& hi) const
  {
    lo = variable_lower_bounds;
    hi = variable_upper_bounds;
    if (lo.empty()) lo.assign(objective_coefficients.size(), 0.0);
    if (hi.empty()) hi.assign(objective_coefficients.size(), std::numeric_limits<double>::infinity());
  }
};

// MPS ingest (cuopt_b200/csrc/mps_reader.cpp).  Throws lp_error(MpsFileError) when the
// file cannot be opened and lp_error(MpsParseError) when it is malformed.
lp_problem_t read_mps(const std::string& path, bool fixed_format = false);

}  // namespace cuopt_b200

// `user_problem_file` and `solution_file` of the parameter registry (constants.h): what the reference's solve_lp does
// around the solve (cpp/src/linear_programming/solve.cu:586-601).  Host code only.
//
// write_problem_as_mps follows cpp/src/mip/problem/write_mps.cu:32-166 section by section, including its choices:
//   * names default to R<i> / C<j>; coefficients at max_digits10; the objective is written with the USER's sign
//   * a row is 'E' when lc == uc, 'G' when uc is infinite, otherwise 'L'
//   * RHS = uc when lc is infinite, lc otherwise (so a RANGED row is written as 'L' with rhs = lc and
//     RANGES = uc - lc, exactly as the reference does — a reader applying the MPS rule for 'L' rows reconstructs
//     [lc - (uc - lc), lc] from that; kept for output parity and flagged here)
//   * BOUNDS: FR when both infinite; otherwise LO (MI for -inf) when lb != 0 or c_j == 0 or the variable is integer,
//     and UP when ub is finite
// write_solution_file follows cpp/src/math_optimization/solution_writer.cu:26-50 and
// linear_programming/solver_solution.cu:369-386 ("Infeasible" for anything but Optimal / PrimalFeasible).
#include "lp_problem.hpp"
#include "pdlp_types.hpp"

#include <cmath>
#include <fstream>
#include <iomanip>
#include <limits>
#include <string>
#include <vector>

namespace cuopt_b200 {

bool write_problem_as_mps(const lp_problem_t& p, const std::string& path)
{
  std::ofstream f(path);
  if (!f.is_open()) return false;
  const double inf = std::numeric_limits<double>::infinity();
  const int m = p.n_constraints, n = p.n_variables;
  hvec<double> lc, uc, lb, ub;
  p.row_bounds(lc, uc);
  p.variable_bounds(lb, ub);
  // column-major copy of A (the reference walks its transposed CSR)
  std::vector<int> coff((size_t)n + 1, 0), crow(p.A_indices.size());
  std::vector<double> cval(p.A_values.size());
  for (int j : p.A_indices) coff[(size_t)j + 1]++;
  for (int j = 0; j < n; ++j) coff[(size_t)j + 1] += coff[j];
  {
    std::vector<int> pos(coff.begin(), coff.end() - 1);
    for (int i = 0; i < m; ++i)
      for (int q = p.A_offsets[i]; q < p.A_offsets[i + 1]; ++q) {
        const int dst = pos[p.A_indices[q]]++;
        crow[dst]     = i;
        cval[dst]     = p.A_values[q];
      }
  }
  auto row_name = [&](int i) { return (size_t)i < p.row_names.size() ? p.row_names[i] : "R" + std::to_string(i); };
  auto col_name = [&](int j) { return (size_t)j < p.variable_names.size() ? p.variable_names[j] : "C" + std::to_string(j); };
  auto is_int   = [&](int j) { return (size_t)j < p.variable_types.size() && p.variable_types[j] != 'C'; };
  const std::string obj = p.objective_name.empty() ? "OBJ" : p.objective_name;

  f << std::setprecision(std::numeric_limits<double>::max_digits10);
  f << "NAME          " << p.problem_name << "\n";
  if (p.maximize) f << "OBJSENSE\n MAXIMIZE\n";
  f << "ROWS\n";
  f << " N  " << obj << "\n";
  for (int i = 0; i < m; ++i) {
    char type = 'L';
    if (lc[i] == uc[i]) type = 'E';
    else if (std::isinf(uc[i])) type = 'G';
    f << " " << type << "  " << row_name(i) << "\n";
  }
  f << "COLUMNS\n";
  bool in_integer_section = false;
  for (int j = 0; j < n; ++j) {
    if (is_int(j) && !in_integer_section) {
      f << "    MARK0001  'MARKER'                 'INTORG'\n";
      in_integer_section = true;
    }
    const double cj = p.objective_coefficients[j];  // the user's sign (the reference un-negates its internal copy)
    if (cj != 0.0) f << "    " << col_name(j) << " " << obj << " " << cj << "\n";
    for (int q = coff[j]; q < coff[(size_t)j + 1]; ++q)
      f << "    " << col_name(j) << " " << row_name(crow[q]) << " " << cval[q] << "\n";
    if (is_int(j) && in_integer_section && (j == n - 1 || !is_int(j + 1))) {
      f << "    MARK0001  'MARKER'                 'INTEND'\n";
      in_integer_section = false;
    }
  }
  f << "RHS\n";
  for (int i = 0; i < m; ++i) {
    const double rhs = std::isinf(lc[i]) ? uc[i] : lc[i];
    if (std::isfinite(rhs) && rhs != 0.0) f << "    RHS1      " << row_name(i) << " " << rhs << "\n";
  }
  bool has_ranges = false;
  for (int i = 0; i < m; ++i) {
    if (lc[i] != -inf && uc[i] != inf && lc[i] != uc[i]) {
      if (!has_ranges) {
        f << "RANGES\n";
        has_ranges = true;
      }
      f << "    RNG1      " << row_name(i) << " " << (uc[i] - lc[i]) << "\n";
    }
  }
  f << "BOUNDS\n";
  for (int j = 0; j < n; ++j) {
    // the reference tests its INTERNAL objective (negated when maximising) against zero: same truth value
    const double cj = p.objective_coefficients[j];
    if (lb[j] == -inf && ub[j] == inf) {
      f << " FR BOUND1    " << col_name(j) << "\n";
    } else {
      if (lb[j] != 0.0 || cj == 0.0 || is_int(j)) {
        if (lb[j] == -inf) f << " MI BOUND1    " << col_name(j) << "\n";
        else f << " LO BOUND1    " << col_name(j) << " " << lb[j] << "\n";
      }
      if (ub[j] != inf) f << " UP BOUND1    " << col_name(j) << " " << ub[j] << "\n";
    }
  }
  f << "ENDATA\n";
  return true;
}

bool write_solution_file(const lp_problem_t& p, const lp_solution_t& s, const std::string& path)
{
  std::ofstream f(path);
  if (!f.is_open()) return false;
  std::string status = termination_status_string(s.termination_status);
  if (s.termination_status != termination_status_t::Optimal && s.termination_status != termination_status_t::PrimalFeasible)
    status = "Infeasible";
  f.precision(std::numeric_limits<double>::max_digits10 + 1);
  f << "# Status: " << status << std::endl;
  if (status != "Infeasible") {
    f << "# Objective value: " << s.stats.primal_objective << std::endl;
    // one line per NAMED variable (the reference iterates over var_names_)
    for (size_t j = 0; j < p.variable_names.size() && j < s.primal.size(); ++j)
      f << p.variable_names[j] << " " << s.primal[j] << std::endl;
  }
  return true;
}

}  // namespace cuopt_b200

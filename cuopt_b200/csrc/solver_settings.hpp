// Named-parameter registry behind cuOptSet*/Get*Parameter.
// Same names, defaults and admissible ranges as the reference's solver_settings_t
// (cpp/src/math_optimization/solver_settings.cu:63-125); MIP-only parameters are stored so that
// client code setting them keeps working, but nothing reads them in this LP-only build.
#pragma once

#include "pdlp_types.hpp"

#include <stdexcept>
#include <string>
#include <vector>

namespace cuopt_b200 {

class solver_settings_t {
 public:
  solver_settings_t();

  // all throw std::invalid_argument on unknown name / unparsable / out-of-range value
  void set_from_string(const std::string& name, const std::string& value);
  std::string get_as_string(const std::string& name) const;
  void set_int(const std::string& name, int value);    // also reaches bool parameters (cuopt_c.cpp:493-503)
  int get_int(const std::string& name) const;          // idem (cuopt_c.cpp:520-532)
  void set_float(const std::string& name, double value);
  double get_float(const std::string& name) const;

  const pdlp_settings_t& pdlp() const { return pdlp_; }
  // extension state that is not a named parameter (warm start, see cuopt_b200_ext.h)
  void set_warm_start(std::shared_ptr<const pdlp_warm_start_t> w) { pdlp_.warm_start = std::move(w); }
  void set_capture_warm_start(bool on) { pdlp_.capture_warm_start = on; }

 private:
  template <typename T>
  struct param_t {
    std::string name;
    T* ptr;
    T lo, hi;
  };
  pdlp_settings_t pdlp_;
  // MIP-only storage
  double mip_time_limit_, mip_abs_tol_, mip_rel_tol_, mip_int_tol_, mip_abs_gap_, mip_rel_gap_;
  int mip_num_cpu_threads_;
  bool mip_scaling_, mip_heuristics_only_, mip_log_to_console_;
  std::string mip_log_file_, mip_sol_file_, mip_user_problem_file_;

  std::vector<param_t<double>> floats_;
  std::vector<param_t<int>> ints_;
  std::vector<param_t<bool>> bools_;
  std::vector<param_t<std::string>> strings_;
};

}  // namespace cuopt_b200

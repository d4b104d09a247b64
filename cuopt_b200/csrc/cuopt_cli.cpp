// cuopt_cli — command-line runner over the C ABI of this library.
//
// Mirrors the reference's two runners for the LP path:
//   cpp/cuopt_cli.cpp                                       cuopt_cli <file.mps> [--<parameter-name> value]... [--relaxation]
//                                                           every solver parameter is an option, '_' written as '-'
//   benchmarks/linear_programming/cuopt/run_pdlp.cu:39-160  --path, --time-limit, --iteration-limit, --optimality-tolerance,
//                                                           --pdlp-solver-mode Stable2|Methodical1|Fast1|Stable1, --method,
//                                                           --crossover, --solution-path
// and adds the batch mode the reference reaches through call_batch_solve (utilities/cython_solve.cu:233-289): several MPS
// files on one command line are solved as independent replicas, dealt round robin to `--gpus N` worker threads, one per
// device (small LPs do not shard: SURVEY.md §8e "replicas only").
//
// Exit code 0 when every file was read and solved (whatever the termination status), 1 otherwise — like the reference.
#include <cuopt_b200/cuopt_b200_ext.h>

#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {

struct options_t {
  std::vector<std::string> files;
  std::vector<std::pair<std::string, std::string>> params;  // parameter name (with '_'), value as given
  bool relaxation = false;
  int gpus        = 1;
  std::string initial_solution;
};

const char* status_name(int s)
{
  switch (s) {
    case CUOPT_TERIMINATION_STATUS_OPTIMAL: return "Optimal";
    case CUOPT_TERIMINATION_STATUS_INFEASIBLE: return "PrimalInfeasible";
    case CUOPT_TERIMINATION_STATUS_UNBOUNDED: return "DualInfeasible";
    case CUOPT_TERIMINATION_STATUS_ITERATION_LIMIT: return "IterationLimit";
    case CUOPT_TERIMINATION_STATUS_TIME_LIMIT: return "TimeLimit";
    case CUOPT_TERIMINATION_STATUS_NUMERICAL_ERROR: return "NumericalError";
    case CUOPT_TERIMINATION_STATUS_PRIMAL_FEASIBLE: return "PrimalFeasible";
    default: return "NoTermination";
  }
}

void usage()
{
  std::fprintf(stderr,
               "usage: cuopt_cli <file.mps> [more.mps ...] [options]\n"
               "  --<parameter-name> <value>   any solver parameter of constants.h, '_' written as '-'\n"
               "                               (--time-limit 10 --optimality-tolerance 1e-6 --pdlp-solver-mode 1 --method 1 ...)\n"
               "  --pdlp-solver-mode also takes Stable1 | Stable2 | Methodical1 | Fast1\n"
               "  --path <file.mps>            same as a positional file (run_pdlp.cu)\n"
               "  --solution-path <file>       same as --solution-file\n"
               "  --relaxation                 solve the LP relaxation of a problem with integer variables\n"
               "  --gpus <N>                   batch mode: deal the files round robin to N devices (default 1)\n"
               "  --version\n");
}

bool parse(int argc, char** argv, options_t& o)
{
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    if (a == "--help" || a == "-h") {
      usage();
      std::exit(0);
    }
    if (a == "--version") {
      std::printf("%s\n", cuOptB200Version());
      std::exit(0);
    }
    if (a.rfind("--", 0) != 0) {
      o.files.push_back(a);
      continue;
    }
    if (a == "--relaxation") {
      o.relaxation = true;
      continue;
    }
    if (i + 1 >= argc) {
      std::fprintf(stderr, "cuopt_cli: option %s needs a value\n", a.c_str());
      return false;
    }
    std::string v = argv[++i];
    if (a == "--path") o.files.push_back(v);
    else if (a == "--gpus") o.gpus = std::max(1, std::atoi(v.c_str()));
    else if (a == "--initial-solution") o.initial_solution = v;
    else {
      std::string name = a.substr(2);
      std::replace(name.begin(), name.end(), '-', '_');
      if (name == "solution_path") name = CUOPT_SOLUTION_FILE;
      if (name == CUOPT_PDLP_SOLVER_MODE) {
        static const std::map<std::string, std::string> modes = {
          {"Stable1", "0"}, {"Stable2", "1"}, {"Methodical1", "2"}, {"Fast1", "3"}};
        auto it = modes.find(v);
        if (it != modes.end()) v = it->second;
      }
      if (name == "optimality_tolerance") {  // run_pdlp.cu:52-55, :104: set_optimality_tolerance = all six tolerances
        for (const char* t : {CUOPT_ABSOLUTE_DUAL_TOLERANCE, CUOPT_RELATIVE_DUAL_TOLERANCE, CUOPT_ABSOLUTE_PRIMAL_TOLERANCE,
                              CUOPT_RELATIVE_PRIMAL_TOLERANCE, CUOPT_ABSOLUTE_GAP_TOLERANCE, CUOPT_RELATIVE_GAP_TOLERANCE})
          o.params.emplace_back(t, v);
        continue;
      }
      o.params.emplace_back(name, v);
    }
  }
  return true;
}

// the problem with every variable continuous (cuopt_cli.cpp --relaxation; relaxed_lp.cu:53-127 does this inside the MIP code)
bool make_relaxation(cuOptOptimizationProblem p, cuOptOptimizationProblem* out)
{
  cuopt_int_t m = 0, n = 0, nnz = 0, sense = 0;
  cuopt_float_t offset = 0;
  if (cuOptGetNumConstraints(p, &m) || cuOptGetNumVariables(p, &n) || cuOptGetNumNonZeros(p, &nnz) ||
      cuOptGetObjectiveSense(p, &sense) || cuOptGetObjectiveOffset(p, &offset))
    return false;
  std::vector<cuopt_int_t> off((size_t)m + 1), idx((size_t)nnz);
  std::vector<cuopt_float_t> val((size_t)nnz), c((size_t)n), lb((size_t)n), ub((size_t)n), clb((size_t)m), cub((size_t)m);
  if (cuOptGetConstraintMatrix(p, off.data(), idx.data(), val.data()) || cuOptGetObjectiveCoefficients(p, c.data()) ||
      cuOptGetVariableLowerBounds(p, lb.data()) || cuOptGetVariableUpperBounds(p, ub.data()) ||
      cuOptGetConstraintLowerBounds(p, clb.data()) || cuOptGetConstraintUpperBounds(p, cub.data()))
    return false;
  std::vector<char> types((size_t)n, CUOPT_CONTINUOUS);
  return cuOptCreateRangedProblem(m, n, sense, offset, c.data(), off.data(), idx.data(), val.data(), clb.data(), cub.data(),
                                  lb.data(), ub.data(), types.data(), out) == CUOPT_SUCCESS;
}

std::mutex g_print;

int run_file(const std::string& path, const options_t& o, int device, bool quiet_library)
{
  const std::string base = path.substr(path.find_last_of("/\\") + 1);
  cuOptOptimizationProblem problem = nullptr;
  cuOptSolverSettings settings     = nullptr;
  cuOptSolution solution           = nullptr;
  int rc                           = 0;
  auto cleanup = [&]() {
    cuOptDestroySolution(&solution);
    cuOptDestroySolverSettings(&settings);
    cuOptDestroyProblem(&problem);
  };
  const auto t0 = std::chrono::steady_clock::now();
  if (cuOptReadProblem(path.c_str(), &problem) != CUOPT_SUCCESS) {
    std::lock_guard<std::mutex> g(g_print);
    std::fprintf(stderr, "Parsing MPS failed. Exiting! (%s)\n", path.c_str());
    return 1;
  }
  cuopt_int_t is_mip = 0;
  cuOptIsMIP(problem, &is_mip);
  if (is_mip && o.relaxation) {
    cuOptOptimizationProblem relaxed = nullptr;
    if (!make_relaxation(problem, &relaxed)) {
      cleanup();
      return 1;
    }
    cuOptDestroyProblem(&problem);
    problem = relaxed;
  }
  if (cuOptCreateSolverSettings(&settings) != CUOPT_SUCCESS) {
    cleanup();
    return 1;
  }
  if (quiet_library) cuOptSetParameter(settings, CUOPT_LOG_TO_CONSOLE, "false");  // batch mode prints one line per file
  for (const auto& kv : o.params) {
    if (cuOptSetParameter(settings, kv.first.c_str(), kv.second.c_str()) != CUOPT_SUCCESS) {
      std::lock_guard<std::mutex> g(g_print);
      std::fprintf(stderr, "Error: unknown parameter or bad value: --%s %s\n", kv.first.c_str(), kv.second.c_str());
      cleanup();
      return 1;
    }
  }
  const cuopt_int_t status = cuOptSolve(problem, settings, &solution);
  const double wall        = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  std::lock_guard<std::mutex> g(g_print);
  if (status != CUOPT_SUCCESS) {
    char msg[512] = "";
    if (solution) cuOptGetErrorString(solution, msg, sizeof(msg));
    std::fprintf(stderr, "Error: %s (%s)\n", msg, base.c_str());
    rc = 1;
  } else {
    cuopt_int_t term = 0;
    cuOptB200LPStats st;
    cuOptGetTerminationStatus(solution, &term);
    cuOptB200GetLPStats(solution, &st);
    std::printf("%-28s gpu %d  Status: %-16s Objective: %+.8e  Dual: %+.8e  Iterations: %7d  Solve: %.3fs  Total: %.3fs\n",
                base.c_str(), device, status_name(term), st.primal_objective, st.dual_objective, st.number_of_steps_taken,
                st.solve_time, wall);
  }
  cleanup();
  return rc;
}

}  // namespace

int main(int argc, char** argv)
{
  options_t o;
  if (!parse(argc, argv, o)) return 1;
  if (o.files.empty()) {
    usage();
    return 1;
  }
  if (!o.initial_solution.empty()) {
    std::fprintf(stderr, "cuopt_cli: --initial-solution is not supported by this build (PDLP starts from x = y = 0 or from a "
                         "cuOptB200 warm start)\n");
    return 1;
  }
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess) n_dev = 0;
  const int workers = std::max(1, std::min({o.gpus, std::max(n_dev, 1), (int)o.files.size()}));
  const bool batch  = o.files.size() > 1;
  std::atomic<int> failures{0};
  auto work = [&](int w) {
    if (n_dev > 0) cudaSetDevice(w % n_dev);
    for (size_t i = w; i < o.files.size(); i += workers) failures += run_file(o.files[i], o, w, batch);
  };
  if (workers == 1) {
    work(0);
  } else {
    std::vector<std::thread> pool;
    for (int w = 0; w < workers; ++w) pool.emplace_back(work, w);
    for (auto& t : pool) t.join();
  }
  return failures.load() ? 1 : 0;
}

"""Small driver for ncu: builds configs[1] (or a smaller LP with --rows) and runs the three PDHG kernels in situ."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cuopt_b200 import capi, lpgen
ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=1_000_000)
ap.add_argument("--cols", type=int, default=1_000_000)
ap.add_argument("--warmup", type=int, default=45)
ap.add_argument("--reps", type=int, default=10)
a = ap.parse_args()
lp = lpgen.sparse_lp(a.rows, a.cols, 8, seed=1234)
p = capi.Problem.create_ranged(lp.offsets, lp.indices, lp.values, lp.con_lb, lp.con_ub, lp.c, lp.var_lb, lp.var_ub)
s = capi.Settings(method=1, log_to_console=False)
s.set("optimality_tolerance", 0.0)
prof = capi.Solver(p, s).profile_kernels(a.warmup, a.reps)
print("ms", prof.ms_primal_step, prof.ms_dual_step, prof.ms_transpose_step, prof.ms_iteration)

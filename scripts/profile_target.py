"""Small driver for ncu: builds a bench workload (--workload c2 | c4 | c3, or --rows/--cols) and runs the three PDHG
kernels in situ (cuOptB200SolverProfileKernels)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cuopt_b200 import capi, lpgen
ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c2", choices=["c2", "c4", "c3"])
ap.add_argument("--rows", type=int, default=0)
ap.add_argument("--cols", type=int, default=0)
ap.add_argument("--warmup", type=int, default=45)
ap.add_argument("--reps", type=int, default=10)
a = ap.parse_args()
if a.workload == "c3":
    lp = lpgen.multicommodity(nodes=9000, arcs=27000, commodities=11, seed=1234)
else:
    size = {"c2": 1_000_000, "c4": 10_000_000}[a.workload]
    lp = lpgen.sparse_lp(a.rows or size, a.cols or size, 8, seed=1234)
p = capi.Problem.create_ranged(lp.offsets, lp.indices, lp.values, lp.con_lb, lp.con_ub, lp.c, lp.var_lb, lp.var_ub)
s = capi.Settings(method=1, log_to_console=False)
s.set("optimality_tolerance", 0.0)
prof = capi.Solver(p, s).profile_kernels(a.warmup, a.reps)
print("ms", prof.ms_primal_step, prof.ms_dual_step, prof.ms_transpose_step, prof.ms_iteration)

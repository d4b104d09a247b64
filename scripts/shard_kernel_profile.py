#!/usr/bin/env python
"""Per-rank kernel times of an N-way sharded solve, measured on ONE GPU: builds the workload, keeps rank 0's block of
rows (all columns) as a problem of its own and profiles the kernels on it.  What it shows: K2 on the shard (gathers
from the full-length xbar), and the payload-free partial product A_g^T y_g with 32-row blocks vs the wide schedule.
usage: python scripts/shard_kernel_profile.py [--workload c4] [--world 8]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cuopt_b200 import capi, lpgen  # noqa: E402
from cuopt_b200 import dist as cdist  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c4", choices=["c2", "c4"])
ap.add_argument("--world", type=int, default=8)
ap.add_argument("--reps", type=int, default=50)
a = ap.parse_args()
size = {"c2": 1_000_000, "c4": 10_000_000}[a.workload]
lp = lpgen.sparse_lp(size, size, 8, seed=1234)
s = capi.Settings(method=1, log_to_console=False)
s.set("optimality_tolerance", 0.0)
for world in sorted({1, a.world}):
    p, (r0, r1) = cdist.local_problem(lp, 0, world) if world > 1 else (
        capi.Problem.create_ranged(lp.offsets, lp.indices, lp.values, lp.con_lb, lp.con_ub, lp.c, lp.var_lb, lp.var_ub),
        (0, lp.m))
    prof = capi.Solver(p, s).profile_kernels(40, a.reps)
    print(json.dumps({"world": world, "rows": r1 - r0, "cols": lp.n, "ms_primal_step_full_n": prof.ms_primal_step,
                      "ms_dual_step": prof.ms_dual_step, "ms_transpose_step_fused": prof.ms_transpose_step,
                      "ms_transpose_partial_32row_blocks": prof.ms_transpose_partial,
                      "ms_transpose_partial_wide_blocks": prof.ms_transpose_partial_wide}), flush=True)

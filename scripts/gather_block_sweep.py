#!/usr/bin/env python
"""Per-kernel times of the PDHG attempt as a function of the gather block size (CUOPT_B200_GATHER_BLOCK_BYTES) and of
the L2 hints, on one workload generated once.  usage: python scripts/gather_block_sweep.py [--workload c4]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cuopt_b200 import capi, lpgen  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c4", choices=["c2", "c4"])
ap.add_argument("--mb", default="0,12,16,20,27,40")
ap.add_argument("--reps", type=int, default=30)
a = ap.parse_args()
size = {"c2": 1_000_000, "c4": 10_000_000}[a.workload]
lp = lpgen.sparse_lp(size, size, 8, seed=1234)
p = capi.Problem.create_ranged(lp.offsets, lp.indices, lp.values, lp.con_lb, lp.con_ub, lp.c, lp.var_lb, lp.var_ub)
s = capi.Settings(method=1, log_to_console=False)
s.set("optimality_tolerance", 0.0)
for mb in [float(v) for v in a.mb.split(",")]:
    for hints in ((1,) if mb else (1, 0)):
        os.environ["CUOPT_B200_GATHER_BLOCK_BYTES"] = str(int(mb * (1 << 20)))
        os.environ["CUOPT_B200_L2_HINTS"] = str(hints)
        prof = capi.Solver(p, s).profile_kernels(45, a.reps)
        print(json.dumps({"block_mb": mb, "l2_hints": hints, "us_k1": round(prof.ms_primal_step * 1e3, 1),
                          "us_k2": round(prof.ms_dual_step * 1e3, 1), "us_k3": round(prof.ms_transpose_step * 1e3, 1),
                          "us_attempt_in_batch": round(prof.ms_iteration * 1e3, 1)}), flush=True)

#!/usr/bin/env python
"""Where does the end-to-end time of one C-ABI solve go?  (host buffers -> cuOptCreateRangedProblem -> cuOptSolve -> getters)
   CUOPT_B200_TRACE=1 python scripts/exp_e2e_trace.py c4 200"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cuopt_b200 import capi, lpgen  # noqa: E402

wl, iters = sys.argv[1], int(sys.argv[2])
n = {"c4": 10_000_000, "c2": 1_000_000}[wl]
lp = lpgen.sparse_lp(n, n, 8, seed=1234)
s = capi.Settings(method=capi.CUOPT_METHOD_PDLP, log_to_console=False, iteration_limit=iters)
s.set("optimality_tolerance", 0.0)
for rep in range(3):
    t0 = time.perf_counter()
    p = capi.Problem.create_ranged(lp.offsets, lp.indices, lp.values, lp.con_lb, lp.con_ub, lp.c, lp.var_lb, lp.var_ub)
    t1 = time.perf_counter()
    sol = capi.solve(p, s)
    t2 = time.perf_counter()
    x = sol.primal(); y = sol.dual()
    t3 = time.perf_counter()
    st = sol.stats()
    print(json.dumps({"rep": rep, "create_s": t1 - t0, "solve_s": t2 - t1, "getters_s": t3 - t2, "setup_seconds": st.setup_seconds,
                      "pdhg_loop_seconds": st.pdhg_loop_seconds, "termination_seconds": st.termination_seconds,
                      "solve_time": st.solve_time, "iterations": st.number_of_steps_taken}), flush=True)
    del sol, p

#!/usr/bin/env python
"""How long does PDLP need to bring configs[1] / configs[3] to optimality_tolerance (gap AND residuals, relative)?
   python scripts/exp_time_to_gap.py c2 1e-6 600000 300"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cuopt_b200 import capi, lpgen  # noqa: E402

wl, tol, itlim, tlim = sys.argv[1], float(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4])
n = {"c4": 10_000_000, "c2": 1_000_000, "c2s": 250_000}[wl]
lp = lpgen.sparse_lp(n, n, 8, seed=1234)
p = capi.Problem.create_ranged(lp.offsets, lp.indices, lp.values, lp.con_lb, lp.con_ub, lp.c, lp.var_lb, lp.var_ub)
s = capi.Settings(method=capi.CUOPT_METHOD_PDLP, log_to_console=False, iteration_limit=itlim, time_limit=tlim)
s.set("optimality_tolerance", tol)
g = capi.Solver(p, s)
g.initialise()
t0 = time.time()
done = False
log = []
chunk = 4000 if wl != "c4" else 2000
while not done:
    done = g.advance(chunk)
    k = int(g.scalar("k_total"))
    if len(log) % 10 == 0 or done:
        print(json.dumps({"k": k, "t": round(time.time() - t0, 2), "step": g.scalar("step_size"), "w": g.scalar("primal_weight"),
                          "restarts": g.scalar("n_restarts")}), flush=True)
    log.append(k)
sol = g.solution()
st = sol.stats()
opt = lp.optimal_objective
print(json.dumps({"workload": wl, "tol": tol, "status": sol.termination_reason, "iterations": st.number_of_steps_taken,
                  "seconds": round(time.time() - t0, 2), "primal_objective": st.primal_objective,
                  "dual_objective": st.dual_objective, "planted": opt,
                  "rel_err_primal": abs(st.primal_objective - opt) / abs(opt), "rel_err_dual": abs(st.dual_objective - opt) / abs(opt),
                  "relative_gap": st.relative_gap, "rel_primal_res": st.l2_relative_primal_residual,
                  "rel_dual_res": st.l2_relative_dual_residual, "restarts": st.n_restarts}))

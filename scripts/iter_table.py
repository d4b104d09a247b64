"""GPU vs oracle iteration counts on the on-disk instances (prints a markdown table)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import problem_arrays
from test_gpu_parity import lp_relaxation, solve_capi
from oracle import pdlp_oracle as po
print("| instance | tol | GPU its | oracle its | GPU obj | oracle obj |")
print("|---|---|---|---|---|---|")
for rel in ["linear_programming/afiro_original.mps", "mip/50v-10-free-bound.mps", "mip/neos5-free-bound.mps",
            "mip/sudoku.mps", "mip/cod105_max.mps", "mip/sample.mps", "mip/bb_optimality.mps"]:
    for tol in (1e-4, 1e-6, 1e-8):
        p = lp_relaxation(rel)
        sol = solve_capi(p, tol=tol, iteration_limit=400000)
        a = problem_arrays(p)
        o = po.Oracle(a["offsets"], a["indices"], a["values"], a["c"], a["var_lb"], a["var_ub"], a["con_lb"], a["con_ub"],
                      maximize=a["maximize"], objective_offset=a["objective_offset"], tol=tol, iteration_limit=400000)
        r = o.solve()
        st = sol.stats()
        print(f"| {rel} | {tol:g} | {st.number_of_steps_taken} | {r['iterations']} | {st.primal_objective:.10g} | {r['primal_objective']:.10g} |")

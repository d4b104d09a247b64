#!/usr/bin/env python
"""tests/golden/c3_reference_simplex.json: the REFERENCE'S OWN CPU dual simplex (cpp/src/dual_simplex, compiled into
oracle/_ref) on the pds-shaped multicommodity LPs of configs[2] — optimal objective, iterations and time on one host
core.  ~1 minute.  Needs /root/reference + `make -C oracle ref`; the output is committed."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cuopt_b200 import lpgen  # noqa: E402
from oracle import ref_cpu  # noqa: E402

out = {"generator": "cuopt_b200.lpgen.multicommodity(nodes, arcs, commodities=11, seed=1234)",
       "solver": "reference dual_simplex::solve_linear_program via oracle/ref_driver.cpp, 1 thread", "cases": []}
for nodes, arcs in ((1200, 3600), (3000, 9000), (9000, 27000)):
    lp = lpgen.multicommodity(nodes=nodes, arcs=arcs, commodities=11, seed=1234)
    t = time.time()
    r = ref_cpu.dual_simplex(lp.offsets, lp.indices, lp.values, lp.con_lb, lp.con_ub, lp.c, lp.var_lb, lp.var_ub,
                             time_limit=600.0)
    out["cases"].append({"nodes": nodes, "arcs": arcs, "rows": lp.m, "cols": lp.n, "nnz": lp.nnz,
                         "status": r["status"], "objective": r["objective"], "iterations": r["iterations"],
                         "seconds": round(r["seconds"], 3), "host_cores_used": 1, "host_cores": os.cpu_count()})
    print(out["cases"][-1], time.time() - t, flush=True)
with open(os.path.join(ROOT, "tests", "golden", "c3_reference_simplex.json"), "w") as f:
    json.dump(out, f, indent=1)

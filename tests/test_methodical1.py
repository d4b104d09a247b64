"""Methodical1 (trust-region restart, pdlp_restart_strategy.cu:278-364, :983-1678) on the CUDA path
(cuopt_b200/csrc/trust_region.cuh), against the oracle restatement that is pinned to the reference's tests:
  * the reference's test_very_low_tolerance (test_lp_solver.py:101-121),
  * the iterates across the first trust-region restarts against the oracle,
  * the objectives of the golden instances (the oracle's are pinned to the reference's dual simplex).
First GPU run: round 2 (gpurun_out r2a); the environment gate of round 1 is gone."""
import numpy as np
import pytest

from conftest import mps_path
from cuopt_b200 import capi
from oracle import pdlp_oracle as po
from test_gpu_parity import make_pair, rel_err

pytestmark = pytest.mark.gpu

# 6 x 64 steps crossing trust-region restarts (sort + prefix sums + bisection feed the restart decision).  The iterates are
# compared with the sequential oracle at every checkpoint; the difference grows along the trajectory: measured on the B200
# at the sixth checkpoint 1.6e-7 with the round-1 SpMV core and 1.06e-6 with the block-interleaved one (different summation
# association inside rows that span lanes).  TRAJECTORY (1e-7) is for 120 steps; the bound here is 1e-5 after 384.
LONG_TRAJECTORY = 1e-5


def test_very_low_tolerance_afiro():
    p = capi.Problem.read(mps_path("linear_programming/afiro_original.mps"))
    s = capi.Settings(method=capi.CUOPT_METHOD_PDLP, log_to_console=False, pdlp_solver_mode=2)
    s.set("optimality_tolerance", 1e-12)
    sol = capi.solve(p, s)
    assert sol.return_code == 0, sol.error_string
    assert sol.termination_status == 1
    assert sol.stats().primal_objective == pytest.approx(-464.7531)
    assert sol.stats().n_restarts >= 1


def test_iterates_across_trust_region_restarts_match_the_oracle():
    g, o, _ = make_pair(capi.Problem.read(mps_path("linear_programming/afiro_original.mps")), mode=2, tol=1e-12)
    g.initialise(); o.initialise()
    for _ in range(6):  # major iterations every 64 steps
        g.advance(64); o.run(64)
        for name in ("x", "y", "aty", "sum_x", "sum_y", "x_last_restart", "y_last_restart"):
            assert rel_err(g.vector(name), o.vector(name)) <= LONG_TRAJECTORY, name
        for name in ("step_size", "primal_weight", "its_since_restart", "n_restarts"):
            assert g.scalar(name) == pytest.approx(o.scalar(name), rel=LONG_TRAJECTORY), name


@pytest.mark.parametrize("rel", ["mip/sudoku.mps", "mip/sample.mps", "mip/bb_optimality.mps"])
def test_objective_against_the_oracle(rel):
    from test_gpu_parity import lp_relaxation
    from conftest import problem_arrays
    p = lp_relaxation(rel)
    a = problem_arrays(p)
    o = po.Oracle(a["offsets"], a["indices"], a["values"], a["c"], a["var_lb"], a["var_ub"], a["con_lb"], a["con_ub"],
                  maximize=a["maximize"], objective_offset=a["objective_offset"], mode=po.METHODICAL1, tol=1e-8,
                  iteration_limit=400000)
    assert o.run(-1) and o.stats().termination_status == 1
    s = capi.Settings(method=capi.CUOPT_METHOD_PDLP, log_to_console=False, pdlp_solver_mode=2)
    s.set("optimality_tolerance", 1e-8)
    sol = capi.solve(p, s)
    assert sol.termination_status == 1
    assert sol.stats().primal_objective == pytest.approx(o.stats().primal_objective, rel=1e-6, abs=1e-6)

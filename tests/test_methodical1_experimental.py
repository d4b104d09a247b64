"""Methodical1 (trust-region restart) on the CUDA path — EXPERIMENTAL.

The device code (cuopt_b200/csrc/trust_region.cuh) was written in round 1 against the oracle restatement after the GPU
budget of the round was spent, so it has not run yet: it is reachable only with CUOPT_B200_EXPERIMENTAL_METHODICAL1=1
(default: pdlp_solver_mode=2 answers CUOPT_VALIDATION_ERROR) and these tests only run with
CUOPT_B200_RUN_EXPERIMENTAL=1.  They are the acceptance tests for switching it on:
  * the reference's test_very_low_tolerance (test_lp_solver.py:101-121),
  * the iterates across the first trust-region restarts against the oracle,
  * the dual-simplex objectives of the golden instances."""
import os

import numpy as np
import pytest

from conftest import mps_path
from cuopt_b200 import capi
from oracle import pdlp_oracle as po
from test_gpu_parity import TRAJECTORY, make_pair, rel_err

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("CUOPT_B200_RUN_EXPERIMENTAL") != "1",
                                 reason="experimental: set CUOPT_B200_RUN_EXPERIMENTAL=1 (see module docstring)")]


@pytest.fixture(autouse=True)
def enable(monkeypatch):
    monkeypatch.setenv("CUOPT_B200_EXPERIMENTAL_METHODICAL1", "1")


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
            assert rel_err(g.vector(name), o.vector(name)) <= TRAJECTORY, name
        for name in ("step_size", "primal_weight", "its_since_restart", "n_restarts"):
            assert g.scalar(name) == pytest.approx(o.scalar(name), rel=1e-7), name


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

"""An LP whose optimum is the starting point x = 0, y = 0 (any LP with c >= 0, l = 0, lc <= 0 <= uc): the very first PDHG
step has zero movement, the step-size rule flags it (adaptive_step_size_strategy.cu:110-116) and the reference still counts
the iteration, retries in the next take_step (pdlp.cu:1191 resets the flag) and answers Optimal at the next termination
test (pdlp.cu:580-583 skips the test while k <= 1).  Round-1 finding: the CUDA path never left k = 1 here."""
import numpy as np
import pytest

from cuopt_b200 import capi
from oracle import pdlp_oracle as po

INF = np.inf
# min x + y  s.t.  x + y <= 10,  x, y >= 0
OFF, IDX, VAL = np.array([0, 2], np.int32), np.array([0, 1], np.int32), np.array([1.0, 1.0])
C, LB, UB, CLB, CUB = np.array([1.0, 1.0]), np.zeros(2), np.full(2, INF), np.array([-INF]), np.array([10.0])


def test_oracle_answers_optimal_after_two_iterations():
    o = po.Oracle(OFF, IDX, VAL, C, LB, UB, CLB, CUB, tol=1e-4)
    r = o.solve()
    assert r["status"] == "Optimal" and r["iterations"] == 2
    assert abs(r["primal_objective"]) <= 1e-12


@pytest.mark.gpu
@pytest.mark.timeout(120)
def test_gpu_answers_optimal_instead_of_hanging():
    p = capi.Problem.create_ranged(OFF, IDX, VAL, CLB, CUB, C, LB, UB)
    s = capi.Settings(method=capi.CUOPT_METHOD_PDLP, log_to_console=False, time_limit=60.0)
    sol = capi.solve(p, s)
    assert sol.return_code == 0, sol.error_string
    assert sol.termination_reason == "Optimal"
    assert sol.stats().number_of_steps_taken == 2
    assert np.allclose(sol.primal(), 0.0) and abs(sol.stats().primal_objective) <= 1e-12

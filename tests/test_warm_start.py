"""Warm start across solves.

Property under test = the reference's own test (cpp/tests/linear_programming/pdlp_test.cu:803-854):
    steps(scratch -> tol2) == steps(scratch -> tol1) + steps(warm start from the tol1 solve -> tol2)
CPU part: the oracle's restatement of pdlp.cu:131-181 / :469-489 / :1074-1136 has the property on every instance, and
the C-ABI warm-start handle round-trips its contents (no GPU needed).  GPU part: the CUDA path has the property, hands
out the same state as the oracle, and continues from the ORACLE's state exactly as the oracle does."""
import numpy as np
import pytest

from conftest import mps_path, problem_arrays
from cuopt_b200 import capi, lpgen
from oracle import pdlp_oracle as po


def instances():
    yield "sparse_300", lpgen.sparse_lp(300, 250, 5, seed=3)
    yield "sparse_2000", lpgen.sparse_lp(2000, 1800, 6, seed=5)
    yield "multicommodity", lpgen.multicommodity(60, 200, 4, seed=2)


def afiro_arrays():
    return problem_arrays(capi.Problem.read(mps_path("linear_programming/afiro_original.mps")))


def oracle_solve(a, tol, warm=None, mode=1):
    o = po.Oracle(a["offsets"], a["indices"], a["values"], a["c"], a["var_lb"], a["var_ub"], a["con_lb"], a["con_ub"],
                  maximize=a.get("maximize", False), objective_offset=a.get("objective_offset", 0.0), mode=mode, tol=tol)
    if warm is not None:
        o.set_warm_start(warm)
    assert o.run(-1)
    return o


def lp_arrays(lp):
    return dict(offsets=lp.offsets, indices=lp.indices, values=lp.values, c=lp.c, var_lb=lp.var_lb, var_ub=lp.var_ub,
                con_lb=lp.con_lb, con_ub=lp.con_ub)


@pytest.mark.parametrize("tols", [(1e-1, 1e-2), (1e-2, 1e-4), (1e-3, 1e-6)])
def test_oracle_warm_start_is_additive(tols):
    t1, t2 = tols
    cases = [(name, lp_arrays(lp)) for name, lp in instances()] + [("afiro", afiro_arrays())]
    for name, a in cases:
        scratch, first = oracle_solve(a, t2), oracle_solve(a, t1)
        cont = oracle_solve(a, t2, first.get_warm_start())
        s, f, c = scratch.stats(), first.stats(), cont.stats()
        assert s.termination_status == f.termination_status == c.termination_status == 1, name
        assert s.number_of_steps_taken == f.number_of_steps_taken + c.number_of_steps_taken, name
        # the continued solve walks the same trajectory: same final point up to the x*d/d round trip of the hand-over
        assert c.primal_objective == pytest.approx(s.primal_objective, rel=1e-9, abs=1e-9), name
        assert np.allclose(cont.vector("solution_x"), scratch.vector("solution_x"), rtol=1e-8, atol=1e-9), name


def test_warm_start_handle_round_trip_and_errors():
    m, n = 5, 7
    rng = np.random.default_rng(1)
    data = {k: rng.normal(size=n if p else m) for k, p in zip(capi.WARM_VECTORS, capi.WARM_IS_PRIMAL)}
    data.update(initial_primal_weight=0.7, initial_step_size=0.03, total_pdlp_iterations=120, total_pdhg_iterations=131,
                last_candidate_kkt_score=1.5, last_restart_kkt_score=2.5, sum_solution_weight=3.25,
                iterations_since_last_restart=17)
    w = capi.WarmStart.create(m, n, data)
    back = w.to_dict()
    for k in capi.WARM_VECTORS:
        assert np.array_equal(back[k], data[k]), k
    for k in capi.WARM_SCALARS:
        assert back[k] == data[k], k
    with pytest.raises(capi.CuOptError):
        w.scalar("no_such_scalar")
    with pytest.raises(capi.CuOptError):
        w.vector("no_such_vector")
    with pytest.raises(ValueError):
        capi.WarmStart.create(m, n, {**data, "current_ATY": np.zeros(n + 1)})
    s = capi.Settings()
    s.set_warm_start(w)       # the settings keep their own reference
    w.close()
    s.set_warm_start(None)
    # NULL arguments
    L = capi.lib()
    assert L.cuOptB200GetWarmStart(None, None) == capi.CUOPT_INVALID_ARGUMENT
    assert L.cuOptB200SetWarmStartCapture(None, 1) == capi.CUOPT_INVALID_ARGUMENT
    assert L.cuOptB200CreateWarmStart(1, 1, None, None, None) == capi.CUOPT_INVALID_ARGUMENT


# ------------------------------------------------------------------------------------------------ GPU
def gpu_solve(p, tol, warm=None, mode=1, capture=True):
    s = capi.Settings(method=capi.CUOPT_METHOD_PDLP, log_to_console=False, pdlp_solver_mode=mode)
    s.set("optimality_tolerance", tol)
    s.capture_warm_start(capture)
    if warm is not None:
        s.set_warm_start(warm)
    sol = capi.solve(p, s)
    assert sol.return_code == 0, sol.error_string
    return sol


def gpu_problem(a):
    return capi.Problem.create_ranged(a["offsets"], a["indices"], a["values"], a["con_lb"], a["con_ub"], a["c"],
                                      a["var_lb"], a["var_ub"], maximize=a.get("maximize", False),
                                      objective_offset=a.get("objective_offset", 0.0))


@pytest.mark.gpu
@pytest.mark.parametrize("tols", [(1e-1, 1e-2), (1e-2, 1e-4), (1e-3, 1e-6)])
def test_gpu_warm_start_is_additive(tols):
    t1, t2 = tols
    cases = [(name, lp_arrays(lp)) for name, lp in instances()] + [("afiro", afiro_arrays())]
    for name, a in cases:
        p = gpu_problem(a)
        scratch, first = gpu_solve(p, t2), gpu_solve(p, t1)
        cont = gpu_solve(p, t2, first.warm_start())
        s, f, c = scratch.stats(), first.stats(), cont.stats()
        assert scratch.termination_status == first.termination_status == cont.termination_status == 1, name
        assert s.number_of_steps_taken == f.number_of_steps_taken + c.number_of_steps_taken, name
        assert c.primal_objective == pytest.approx(s.primal_objective, rel=1e-9, abs=1e-9), name
        assert np.allclose(cont.primal(), scratch.primal(), rtol=1e-8, atol=1e-9), name


@pytest.mark.gpu
def test_gpu_warm_start_state_matches_oracle_and_continues_from_it():
    for name, a in [(n_, lp_arrays(lp)) for n_, lp in instances()] + [("afiro", afiro_arrays())]:
        p = gpu_problem(a)
        o1 = oracle_solve(a, 1e-2)
        g1 = gpu_solve(p, 1e-2)
        wo, wg = o1.get_warm_start(), g1.warm_start().to_dict()
        for k in capi.WARM_INT_SCALARS:
            assert wg[k] == wo[k], (name, k)
        for k in ("initial_primal_weight", "initial_step_size", "last_candidate_kkt_score", "last_restart_kkt_score",
                  "sum_solution_weight"):
            assert wg[k] == pytest.approx(wo[k], rel=1e-7), (name, k)
        for k in capi.WARM_VECTORS:
            scale = max(1.0, float(np.max(np.abs(wo[k]))))
            assert np.max(np.abs(wg[k] - wo[k])) <= 1e-7 * scale, (name, k)
        # hand the ORACLE's state to the CUDA path: it must finish the 1e-4 solve where the oracle does
        o2 = oracle_solve(a, 1e-4, wo)
        g2 = gpu_solve(p, 1e-4, capi.WarmStart.create(len(a["con_lb"]), len(a["c"]), wo))
        assert g2.termination_status == 1
        assert g2.stats().number_of_steps_taken == o2.stats().number_of_steps_taken, name
        # same iteration count; the iterates agree to the trajectory tolerance of test_gpu_parity.py (summation order)
        assert g2.stats().primal_objective == pytest.approx(o2.stats().primal_objective, rel=1e-7, abs=1e-9), name


@pytest.mark.gpu
def test_gpu_warm_start_of_the_wrong_size_is_a_validation_error():
    a, b = lp_arrays(lpgen.sparse_lp(300, 250, 5, seed=3)), lp_arrays(lpgen.sparse_lp(200, 150, 5, seed=4))
    w = gpu_solve(gpu_problem(b), 1e-2).warm_start()
    s = capi.Settings(method=capi.CUOPT_METHOD_PDLP, log_to_console=False)
    s.set_warm_start(w)
    sol = capi.solve(gpu_problem(a), s)
    assert sol.return_code == capi.CUOPT_VALIDATION_ERROR
    assert "warm start" in sol.error_string


@pytest.mark.gpu
def test_solution_without_capture_has_no_warm_start():
    sol = gpu_solve(gpu_problem(lp_arrays(lpgen.sparse_lp(300, 250, 5, seed=3))), 1e-2, capture=False)
    with pytest.raises(capi.CuOptError):
        sol.warm_start()


# ------------------------------------------------------------------- per_constraint_residual (reference pin)
def per_constraint_case():
    """The LP of the reference's per_constraint_test (cpp/tests/linear_programming/pdlp_test.cu:633-716): 3 x 3 identity,
    b = 0, c = 0, iterate x = (0.02, 0.03, 0.1): ||r||_2 = 0.1063 > 0.1 but max_i r_i = 0.1 <= 0.1."""
    a = dict(offsets=np.array([0, 1, 2, 3], np.int32), indices=np.array([0, 1, 2], np.int32), values=np.ones(3),
             c=np.zeros(3), var_lb=np.zeros(3), var_ub=np.full(3, np.inf), con_lb=np.zeros(3), con_ub=np.zeros(3))
    x = np.array([0.02, 0.03, 0.1])
    z = np.zeros(3)
    warm = dict(current_primal_solution=x, current_dual_solution=z, initial_primal_average=x, initial_dual_average=z,
                current_ATY=z, sum_primal_solutions=z, sum_dual_solutions=z, last_restart_duality_gap_primal_solution=z,
                last_restart_duality_gap_dual_solution=z, initial_primal_weight=1.0, initial_step_size=0.1,
                total_pdlp_iterations=40, total_pdhg_iterations=40, last_candidate_kkt_score=1.0,
                last_restart_kkt_score=1.0, sum_solution_weight=0.0, iterations_since_last_restart=0)
    tol = dict(abs_primal_tol=0.1, rel_primal_tol=0.0, abs_dual_tol=0.1, rel_dual_tol=0.0)
    return a, x, warm, tol


def test_oracle_per_constraint_residual_matches_the_reference_test():
    a, x, warm, tol = per_constraint_case()
    for per_constraint, want_status in ((False, 6), (True, 1)):
        o = po.Oracle(a["offsets"], a["indices"], a["values"], a["c"], a["var_lb"], a["var_ub"], a["con_lb"], a["con_ub"],
                      tolerances=tol, per_constraint_residual=per_constraint)
        o.initialise()
        cv = o.convergence(x, np.zeros(3))
        assert cv["l2_primal_residual"] == pytest.approx(0.10630145812734649, rel=1e-14)
        assert int(cv["status"]) == want_status       # EXPECT_TRUE(status != Optimal) / Optimal
        if per_constraint:
            assert o.scalar("linf_relative_primal_residual") == 0.1   # EXPECT_EQ(..., 0.1) in the reference test
    # through the whole solver: started AT that iterate, the per-constraint run stops before taking a step
    for per_constraint in (False, True):
        o = po.Oracle(a["offsets"], a["indices"], a["values"], a["c"], a["var_lb"], a["var_ub"], a["con_lb"], a["con_ub"],
                      tolerances=tol, per_constraint_residual=per_constraint)
        o.set_warm_start(warm)
        assert o.run(-1)
        assert o.stats().termination_status == 1
        assert (o.stats().number_of_steps_taken == 0) == per_constraint


@pytest.mark.gpu
def test_gpu_per_constraint_residual_matches_the_reference_test_and_the_oracle():
    a, x, warm, tol = per_constraint_case()
    p = gpu_problem(a)
    for per_constraint in (False, True):
        o = po.Oracle(a["offsets"], a["indices"], a["values"], a["c"], a["var_lb"], a["var_ub"], a["con_lb"], a["con_ub"],
                      tolerances=tol, per_constraint_residual=per_constraint)
        o.set_warm_start(warm)
        assert o.run(-1)
        s = capi.Settings(method=capi.CUOPT_METHOD_PDLP, log_to_console=False, per_constraint_residual=per_constraint)
        for name, v in (("absolute_primal_tolerance", 0.1), ("relative_primal_tolerance", 0.0),
                        ("absolute_dual_tolerance", 0.1), ("relative_dual_tolerance", 0.0)):
            s.set(name, v)
        s.set_warm_start(capi.WarmStart.create(3, 3, warm))
        sol = capi.solve(p, s)
        assert sol.return_code == 0, sol.error_string
        assert sol.termination_status == 1
        assert sol.stats().number_of_steps_taken == o.stats().number_of_steps_taken
        assert (sol.stats().number_of_steps_taken == 0) == per_constraint
        if per_constraint:
            assert np.array_equal(sol.primal(), x)   # the iterate it was handed, judged optimal row by row

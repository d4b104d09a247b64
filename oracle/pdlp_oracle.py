"""TEST INFRASTRUCTURE — ctypes front end of oracle/pdlp_oracle.cpp (CPU restatement of the reference PDLP).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` leg import this.
Presets restate cpp/src/linear_programming/solve.cu:64-199 (defaults: pdlp_hyper_params.cu:22-80).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libpdlp_oracle.so")
_lib = None


class Hyper(C.Structure):
    _fields_ = [
        ("initial_step_size_scaling", C.c_double),
        ("l_inf_ruiz_iterations", C.c_int),
        ("do_pock_chambolle_scaling", C.c_int),
        ("do_ruiz_scaling", C.c_int),
        ("alpha_pock_chambolle", C.c_double),
        ("artificial_restart_threshold", C.c_double),
        ("compute_initial_step_size_before_scaling", C.c_int),
        ("compute_initial_primal_weight_before_scaling", C.c_int),
        ("initial_primal_weight_c_scaling", C.c_double),
        ("initial_primal_weight_b_scaling", C.c_double),
        ("major_iteration", C.c_int),
        ("min_iteration_restart", C.c_int),
        ("restart_strategy", C.c_int),
        ("never_restart_to_average", C.c_int),
        ("reduction_exponent", C.c_double),
        ("growth_exponent", C.c_double),
        ("primal_weight_update_smoothing", C.c_double),
        ("sufficient_reduction_for_restart", C.c_double),
        ("necessary_reduction_for_restart", C.c_double),
        ("primal_importance", C.c_double),
        ("primal_distance_smoothing", C.c_double),
        ("dual_distance_smoothing", C.c_double),
        ("compute_last_restart_before_new_primal_weight", C.c_int),
        ("artificial_restart_in_main_loop", C.c_int),
        ("rescale_for_restart", C.c_int),
        ("handle_some_primal_gradients_on_finite_bounds_as_residuals", C.c_int),
        ("project_initial_primal", C.c_int),
    ]


class Settings(C.Structure):
    _fields_ = [
        ("abs_dual_tol", C.c_double), ("rel_dual_tol", C.c_double),
        ("abs_primal_tol", C.c_double), ("rel_primal_tol", C.c_double),
        ("abs_gap_tol", C.c_double), ("rel_gap_tol", C.c_double),
        ("iteration_limit", C.c_int), ("time_limit", C.c_double), ("num_threads", C.c_int),
        ("per_constraint_residual", C.c_int), ("detect_infeasibility", C.c_int), ("strict_infeasibility", C.c_int),
        ("primal_infeasible_tol", C.c_double), ("dual_infeasible_tol", C.c_double),
        ("save_best_primal_so_far", C.c_int),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("termination_status", C.c_int), ("number_of_steps_taken", C.c_int),
        ("total_number_of_attempted_steps", C.c_int),
        ("primal_objective", C.c_double), ("dual_objective", C.c_double), ("gap", C.c_double),
        ("relative_gap", C.c_double),
        ("l2_primal_residual", C.c_double), ("l2_relative_primal_residual", C.c_double),
        ("l2_dual_residual", C.c_double), ("l2_relative_dual_residual", C.c_double),
        ("step_size", C.c_double), ("primal_weight", C.c_double),
        ("n_restarts", C.c_int), ("n_major", C.c_int), ("solve_seconds", C.c_double),
        ("solution_is_average", C.c_int),
    ]


STABLE1, STABLE2, METHODICAL1, FAST1 = 0, 1, 2, 3
_STABLE2 = dict(  # solve.cu:99-131
    initial_step_size_scaling=1.0, l_inf_ruiz_iterations=10, do_pock_chambolle_scaling=1, do_ruiz_scaling=1,
    alpha_pock_chambolle=1.0, artificial_restart_threshold=0.36, compute_initial_step_size_before_scaling=0,
    compute_initial_primal_weight_before_scaling=0, initial_primal_weight_c_scaling=1.0,
    initial_primal_weight_b_scaling=1.0, major_iteration=40, min_iteration_restart=10, restart_strategy=1,
    never_restart_to_average=0, reduction_exponent=0.3, growth_exponent=0.6, primal_weight_update_smoothing=0.5,
    sufficient_reduction_for_restart=0.2, necessary_reduction_for_restart=0.8, primal_importance=1.0,
    primal_distance_smoothing=0.5, dual_distance_smoothing=0.5, compute_last_restart_before_new_primal_weight=1,
    artificial_restart_in_main_loop=0, rescale_for_restart=1,
    handle_some_primal_gradients_on_finite_bounds_as_residuals=0, project_initial_primal=1)
_DELTA = {
    STABLE2: {},
    STABLE1: dict(  # solve.cu:64-95
        initial_step_size_scaling=1.6, l_inf_ruiz_iterations=1, alpha_pock_chambolle=1.3,
        artificial_restart_threshold=0.5, compute_initial_primal_weight_before_scaling=1,
        initial_primal_weight_c_scaling=2.2, initial_primal_weight_b_scaling=4.6, major_iteration=52,
        min_iteration_restart=0, reduction_exponent=0.5, growth_exponent=0.9, primal_weight_update_smoothing=0.3,
        sufficient_reduction_for_restart=0.2, necessary_reduction_for_restart=0.5, primal_importance=1.8,
        primal_distance_smoothing=0.6, dual_distance_smoothing=0.2, compute_last_restart_before_new_primal_weight=0,
        rescale_for_restart=0, handle_some_primal_gradients_on_finite_bounds_as_residuals=1,
        project_initial_primal=0),
    METHODICAL1: dict(  # solve.cu:133-165
        l_inf_ruiz_iterations=5, artificial_restart_threshold=0.5, major_iteration=64, min_iteration_restart=0,
        restart_strategy=2, sufficient_reduction_for_restart=0.1, necessary_reduction_for_restart=0.9,
        rescale_for_restart=0, handle_some_primal_gradients_on_finite_bounds_as_residuals=1,
        project_initial_primal=0),
    FAST1: dict(  # solve.cu:167-199
        initial_step_size_scaling=0.8, l_inf_ruiz_iterations=6, do_ruiz_scaling=0, alpha_pock_chambolle=2.0,
        artificial_restart_threshold=0.3, compute_initial_primal_weight_before_scaling=1,
        initial_primal_weight_c_scaling=1.2, initial_primal_weight_b_scaling=1.2, major_iteration=76,
        min_iteration_restart=6, never_restart_to_average=1, reduction_exponent=0.4, growth_exponent=0.6,
        sufficient_reduction_for_restart=0.3, necessary_reduction_for_restart=0.9, primal_importance=0.8,
        primal_distance_smoothing=0.8, dual_distance_smoothing=0.3, artificial_restart_in_main_loop=1,
        handle_some_primal_gradients_on_finite_bounds_as_residuals=1, project_initial_primal=0),
}


def preset(mode: int = STABLE2) -> Hyper:
    d = dict(_STABLE2)
    d.update(_DELTA[mode])
    return Hyper(**d)


def build(force: bool = False) -> str:
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(
            os.path.join(_HERE, "pdlp_oracle.cpp")):
        subprocess.check_call(["make", "-C", _HERE, "oracle"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.pdlp_oracle_create.restype = C.c_void_p
        _lib.pdlp_oracle_destroy.argtypes = [C.c_void_p]
        _lib.pdlp_oracle_initialise.argtypes = [C.c_void_p]
        _lib.pdlp_oracle_run.argtypes = [C.c_void_p, C.c_int]
        _lib.pdlp_oracle_run.restype = C.c_int
        _lib.pdlp_oracle_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
        _lib.pdlp_oracle_get_vector.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
        _lib.pdlp_oracle_get_vector.restype = C.c_int
        _lib.pdlp_oracle_get_scalar.argtypes = [C.c_void_p, C.c_char_p]
        _lib.pdlp_oracle_get_scalar.restype = C.c_double
        _lib.pdlp_oracle_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        _lib.pdlp_oracle_trace.restype = C.c_int
        _lib.pdlp_oracle_single_attempt.argtypes = [C.c_void_p] + [C.c_void_p] * 3 + [C.c_double] * 2 + [C.c_void_p] * 5
        _lib.pdlp_oracle_convergence.argtypes = [C.c_void_p] + [C.c_void_p] * 4
        _lib.pdlp_oracle_trust_region_bounds.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
        _lib.pdlp_oracle_get_warm_start.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.pdlp_oracle_set_warm_start.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    return _lib


TERMINATION = {0: "NoTermination", 1: "Optimal", 2: "PrimalInfeasible", 3: "DualInfeasible", 4: "IterationLimit",
               5: "TimeLimit", 6: "NumericalError", 7: "PrimalFeasible", 8: "FeasibleFound", 9: "ConcurrentLimit"}


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


WARM_VECTORS = ("current_primal_solution", "current_dual_solution", "initial_primal_average", "initial_dual_average",
                "current_ATY", "sum_primal_solutions", "sum_dual_solutions", "last_restart_duality_gap_primal_solution",
                "last_restart_duality_gap_dual_solution")
WARM_IS_PRIMAL = (True, False, True, False, True, True, False, True, False)
WARM_SCALARS = ("initial_primal_weight", "initial_step_size", "total_pdlp_iterations", "total_pdhg_iterations",
                "last_candidate_kkt_score", "last_restart_kkt_score", "sum_solution_weight",
                "iterations_since_last_restart")
WARM_INT_SCALARS = ("total_pdlp_iterations", "total_pdhg_iterations", "iterations_since_last_restart")


class Oracle:
    """One PDLP solve on the CPU.  Inputs are the ranged LP  min/max c'x + offset, lc <= Ax <= uc, l <= x <= u."""

    def __init__(self, offsets, indices, values, c, var_lb, var_ub, con_lb, con_ub, *, maximize=False,
                 objective_offset=0.0, mode=STABLE2, hyper: Hyper | None = None, tol=1e-4, iteration_limit=2**31 - 1,
                 time_limit=float("inf"), tolerances: dict | None = None, per_constraint_residual=False,
                 num_threads: int = 0, detect_infeasibility=False, strict_infeasibility=False,
                 primal_infeasible_tol=1e-8, dual_infeasible_tol=1e-8, save_best_primal_so_far=False):
        L = lib()
        self.m, self.n = len(con_lb), len(c)
        self._keep = [np.ascontiguousarray(offsets, np.int32), np.ascontiguousarray(indices, np.int32)] + [
            np.ascontiguousarray(a, np.float64) for a in (values, c, var_lb, var_ub, con_lb, con_ub)]
        self.hyper = hyper if hyper is not None else preset(mode)
        t = dict(abs_dual_tol=tol, rel_dual_tol=tol, abs_primal_tol=tol, rel_primal_tol=tol, abs_gap_tol=tol,
                 rel_gap_tol=tol)
        if tolerances:
            t.update(tolerances)
        self.settings = Settings(iteration_limit=int(iteration_limit), time_limit=float(time_limit),
                                 num_threads=int(num_threads), detect_infeasibility=int(bool(detect_infeasibility)),
                                 strict_infeasibility=int(bool(strict_infeasibility)),
                                 primal_infeasible_tol=float(primal_infeasible_tol),
                                 dual_infeasible_tol=float(dual_infeasible_tol),
                                 save_best_primal_so_far=int(bool(save_best_primal_so_far)),
                                 per_constraint_residual=int(bool(per_constraint_residual)), **t)
        k = self._keep
        self.h = C.c_void_p(L.pdlp_oracle_create(
            C.c_int(self.m), C.c_int(self.n), _p(k[0]), _p(k[1]), _p(k[2]), _p(k[3]), _p(k[4]), _p(k[5]), _p(k[6]),
            _p(k[7]), C.c_int(int(maximize)), C.c_double(objective_offset), C.byref(self.hyper),
            C.byref(self.settings)))

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.pdlp_oracle_destroy(self.h)
            self.h = None

    def initialise(self):
        lib().pdlp_oracle_initialise(self.h)

    def run(self, max_steps: int = -1) -> bool:
        return bool(lib().pdlp_oracle_run(self.h, int(max_steps)))

    def stats(self) -> Stats:
        s = Stats()
        lib().pdlp_oracle_stats(self.h, C.byref(s))
        return s

    def trust_region_bounds(self, px, py, radius: float = -1.0):
        """(lower, upper, radius) of bound_optimal_objective at a point of the scaled space (Methodical1 restart)."""
        px = np.ascontiguousarray(px, np.float64); py = np.ascontiguousarray(py, np.float64)
        out = np.zeros(4)
        lib().pdlp_oracle_trust_region_bounds(self.h, _p(px), _p(py), C.c_double(radius), _p(out))
        return float(out[0]), float(out[1]), float(out[2])

    # ---- warm start (pdlp_warm_start_data.hpp:28-72): dict with the header's field names ----
    def get_warm_start(self) -> dict:
        """State to continue from, as the reference hands it out at termination (pdlp.cu:469-489)."""
        vecs = [np.zeros(self.n if p else self.m) for p in WARM_IS_PRIMAL]
        ptrs = (C.c_void_p * 9)(*[v.ctypes.data for v in vecs])
        sc = np.zeros(8)
        lib().pdlp_oracle_get_warm_start(self.h, ptrs, _p(sc))
        out = dict(zip(WARM_VECTORS, vecs))
        out.update({k: (int(v) if k in WARM_INT_SCALARS else float(v)) for k, v in zip(WARM_SCALARS, sc)})
        return out

    def set_warm_start(self, w: dict):
        """Before initialise() / the first run() (pdlp.cu:131-181)."""
        self._warm = [np.ascontiguousarray(w[k], np.float64) for k in WARM_VECTORS]
        ptrs = (C.c_void_p * 9)(*[v.ctypes.data for v in self._warm])
        sc = np.array([float(w[k]) for k in WARM_SCALARS])
        lib().pdlp_oracle_set_warm_start(self.h, ptrs, _p(sc))

    def vector(self, name: str) -> np.ndarray:
        n = lib().pdlp_oracle_get_vector(self.h, name.encode(), None)
        if n < 0:
            raise KeyError(name)
        out = np.zeros(n)
        lib().pdlp_oracle_get_vector(self.h, name.encode(), _p(out))
        return out

    def scalar(self, name: str) -> float:
        return float(lib().pdlp_oracle_get_scalar(self.h, name.encode()))

    def trace(self) -> np.ndarray:
        r = lib().pdlp_oracle_trace(self.h, None, 0)
        out = np.zeros((r, 12))
        if r:
            lib().pdlp_oracle_trace(self.h, _p(out), r)
        return out

    def single_attempt(self, x, y, aty, tau, sigma):
        x, y, aty = (np.ascontiguousarray(a, np.float64) for a in (x, y, aty))
        xn = np.zeros(self.n); xbar = np.zeros(self.n); yn = np.zeros(self.m); atyn = np.zeros(self.n)
        red = np.zeros(3)
        lib().pdlp_oracle_single_attempt(self.h, _p(x), _p(y), _p(aty), C.c_double(tau), C.c_double(sigma), _p(xn),
                                         _p(xbar), _p(yn), _p(atyn), _p(red))
        return dict(x_next=xn, x_bar=xbar, y_next=yn, aty_next=atyn, interaction=red[0], norm_dx2=red[1],
                    norm_dy2=red[2])

    def convergence(self, x_unscaled, y_unscaled):
        x, y = (np.ascontiguousarray(a, np.float64) for a in (x_unscaled, y_unscaled))
        out = np.zeros(8); rc = np.zeros(self.n)
        lib().pdlp_oracle_convergence(self.h, _p(x), _p(y), _p(out), _p(rc))
        keys = ["l2_primal_residual", "l2_dual_residual", "primal_objective", "dual_objective", "gap",
                "abs_objective", "status", "kkt"]
        d = dict(zip(keys, out))
        d["reduced_cost"] = rc
        return d

    def solve(self):
        self.run(-1)
        s = self.stats()
        return dict(status=TERMINATION[s.termination_status], iterations=s.number_of_steps_taken,
                    attempted=s.total_number_of_attempted_steps, primal_objective=s.primal_objective,
                    dual_objective=s.dual_objective, gap=s.gap, l2_primal_residual=s.l2_primal_residual,
                    l2_dual_residual=s.l2_dual_residual, x=self.vector("solution_x"), y=self.vector("solution_y"),
                    reduced_cost=self.vector("solution_rc"), n_restarts=s.n_restarts, n_major=s.n_major,
                    seconds=s.solve_seconds, step_size=s.step_size, primal_weight=s.primal_weight,
                    is_average=bool(s.solution_is_average))

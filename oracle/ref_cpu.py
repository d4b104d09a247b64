"""TEST INFRASTRUCTURE — ctypes access to oracle/_ref/libcuopt_ref_cpu.so.

That library is the reference's own `libmps_parser` and CPU `dual_simplex`
(compiled unchanged from /root/reference by oracle/Makefile) behind the small
C shim oracle/ref_driver.cpp.  Only tests/, scripts/gen_golden.py and
bench.py's cpu_baseline / `--impl reference` leg may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libcuopt_ref_cpu.so")
_lib = None


def available() -> bool:
    return os.path.exists(_LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise FileNotFoundError(f"{_LIB_PATH} not built (run `make -C oracle ref` where /root/reference exists)")
        _lib = C.CDLL(_LIB_PATH)
        _lib.ref_mps_parse.restype = C.c_void_p
        _lib.ref_mps_parse.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int]
        _lib.ref_mps_free.argtypes = [C.c_void_p]
        _lib.ref_mps_dims.argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 4 + [C.POINTER(C.c_double)] * 2
        _lib.ref_mps_arrays.argtypes = [C.c_void_p] + [C.c_void_p] * 11
        _lib.ref_mps_names.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int]
        _lib.ref_mps_names.restype = C.c_int
        _lib.ref_dual_simplex.restype = C.c_int
    return _lib


@dataclass
class MpsModel:
    """Host arrays exactly as the reference parser produced them (mps_data_model_t)."""
    m: int
    n: int
    nnz: int
    maximize: bool
    objective_offset: float
    objective_scaling_factor: float
    offsets: np.ndarray
    indices: np.ndarray
    values: np.ndarray
    rhs: np.ndarray
    c: np.ndarray
    var_lb: np.ndarray
    var_ub: np.ndarray
    con_lb: np.ndarray
    con_ub: np.ndarray
    row_types: bytes
    var_types: bytes
    var_names: list = field(default_factory=list)
    row_names: list = field(default_factory=list)
    problem_name: str = ""
    objective_name: str = ""


def parse_mps(path: str, fixed_format: bool = False) -> MpsModel:
    L = lib()
    err = C.create_string_buffer(1024)
    h = L.ref_mps_parse(path.encode(), int(fixed_format), err, 1024)
    if not h:
        raise ValueError(err.value.decode(errors="replace"))
    try:
        m, n, nnz, mx = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        off, sc = C.c_double(), C.c_double()
        L.ref_mps_dims(h, m, n, nnz, mx, off, sc)
        m, n, nnz = m.value, n.value, nnz.value
        offsets = np.zeros(m + 1, np.int32)
        indices = np.zeros(nnz, np.int32)
        values = np.zeros(nnz, np.float64)
        rhs = np.zeros(m); c = np.zeros(n); lb = np.zeros(n); ub = np.zeros(n)
        clb = np.zeros(m); cub = np.zeros(m)
        rt = C.create_string_buffer(max(m, 1)); vt = C.create_string_buffer(max(n, 1))
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        L.ref_mps_arrays(h, p(offsets), p(indices), p(values), p(rhs), p(c), p(lb), p(ub), p(clb), p(cub),
                         C.cast(rt, C.c_void_p), C.cast(vt, C.c_void_p))

        def names(which):
            ln = L.ref_mps_names(h, which, None, 0)
            buf = C.create_string_buffer(ln + 1)
            L.ref_mps_names(h, which, buf, ln + 1)
            s = buf.value.decode(errors="replace")
            return s.split("\n") if s else []

        pn = names(2) + ["", ""]
        return MpsModel(m, n, nnz, bool(mx.value), off.value, sc.value, offsets, indices, values, rhs, c, lb, ub,
                        clb, cub, rt.raw[:m], vt.raw[:n], names(0) if n else [], names(1) if m else [], pn[0], pn[1])
    finally:
        L.ref_mps_free(h)


SIMPLEX_STATUS = {0: "OPTIMAL", 1: "INFEASIBLE", 2: "UNBOUNDED", 3: "ITERATION_LIMIT", 4: "TIME_LIMIT",
                  5: "NUMERICAL_ISSUES", 6: "CUTOFF", 7: "CONCURRENT_LIMIT", 8: "UNSET"}


def dual_simplex(offsets, indices, values, con_lb, con_ub, c, var_lb, var_ub, *, maximize=False, objective_offset=0.0,
                 time_limit=float("inf"), iteration_limit=2**31 - 1, log=False):
    """Run the reference's CPU dual simplex (cpp/src/dual_simplex) on a ranged LP.

    Mirrors how the reference hands an LP to it (problem_helpers.cuh:128-142 negates c for
    maximisation and flips the objective scaling; translate.hpp:30-100 converts bounds to senses).
    Returns dict(status, objective, iterations, seconds, x, y, z).
    """
    L = lib()
    m, n = len(con_lb), len(c)
    offsets = np.ascontiguousarray(offsets, np.int32); indices = np.ascontiguousarray(indices, np.int32)
    values = np.ascontiguousarray(values, np.float64)
    cc = np.ascontiguousarray(c, np.float64)
    scale = 1.0
    if maximize:
        cc = -cc
        scale = -1.0
    arrs = [np.ascontiguousarray(a, np.float64) for a in (con_lb, con_ub, cc, var_lb, var_ub)]
    x = np.zeros(n); y = np.zeros(m); z = np.zeros(n)
    obj = C.c_double(); its = C.c_int(); st = C.c_int(); secs = C.c_double()
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
    rc = L.ref_dual_simplex(C.c_int(m), C.c_int(n), ip(offsets), ip(indices), dp(values), dp(arrs[0]), dp(arrs[1]),
                            dp(arrs[2]), dp(arrs[3]), dp(arrs[4]), C.c_double(scale), C.c_double(objective_offset),
                            C.c_double(time_limit), C.c_int(iteration_limit), C.c_int(int(log)), dp(x), dp(y), dp(z),
                            C.byref(obj), C.byref(its), C.byref(st), C.byref(secs))
    if rc != 0:
        raise RuntimeError("reference dual simplex threw")
    return dict(status=SIMPLEX_STATUS.get(st.value, str(st.value)), objective=obj.value, iterations=its.value,
                seconds=secs.value, x=x, y=y, z=z)


def dual_simplex_mps(path: str, **kw):
    mdl = parse_mps(path)
    return dual_simplex(mdl.offsets, mdl.indices, mdl.values, mdl.con_lb, mdl.con_ub, mdl.c, mdl.var_lb, mdl.var_ub,
                        maximize=mdl.maximize, objective_offset=mdl.objective_offset, **kw)

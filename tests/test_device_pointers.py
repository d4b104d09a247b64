"""The C ABI takes host OR device pointers for every array (cuopt_c.cpp:110-135 copies inputs with raft::copy, :261-266
the getter outputs): a problem created from device buffers, solved, and read back into device buffers must equal the same
through host buffers."""
import ctypes as C

import numpy as np
import pytest

from cuopt_b200 import capi, lpgen

pytestmark = pytest.mark.gpu


def test_problem_from_device_buffers_and_getters_into_device_buffers():
    import torch
    lp = lpgen.sparse_lp(4000, 3000, 6, seed=3)
    capi.lib()
    L = C.CDLL(capi.lib_path())  # a second handle: raw-pointer prototypes here, the numpy-typed ones of capi stay intact
    dev = torch.device("cuda", 0)
    t = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in
         dict(off=lp.offsets.astype(np.int32), idx=lp.indices.astype(np.int32), val=lp.values, clb=lp.con_lb, cub=lp.con_ub,
              c=lp.c, lb=lp.var_lb, ub=lp.var_ub).items()}
    vt = torch.full((lp.n,), ord("C"), dtype=torch.uint8, device=dev)
    ptr = lambda x: C.c_void_p(x.data_ptr())
    h = C.c_void_p()
    rc = L.cuOptCreateRangedProblem(C.c_int32(lp.m), C.c_int32(lp.n), C.c_int32(capi.CUOPT_MINIMIZE), C.c_double(0.0),
                                    ptr(t["c"]), ptr(t["off"]), ptr(t["idx"]), ptr(t["val"]), ptr(t["clb"]), ptr(t["cub"]),
                                    ptr(t["lb"]), ptr(t["ub"]), ptr(vt), C.byref(h))
    assert rc == 0
    p_dev = capi.Problem(h)
    p_host = capi.Problem.create_ranged(lp.offsets, lp.indices, lp.values, lp.con_lb, lp.con_ub, lp.c, lp.var_lb, lp.var_ub)
    assert p_dev.num_nonzeros == p_host.num_nonzeros == lp.nnz
    o1, i1, v1 = p_dev.constraint_matrix(); o2, i2, v2 = p_host.constraint_matrix()
    assert np.array_equal(o1, o2) and np.array_equal(i1, i2) and np.array_equal(v1, v2)
    s = capi.Settings(method=capi.CUOPT_METHOD_PDLP, log_to_console=False)
    s.set("optimality_tolerance", 1e-6)
    a, b = capi.solve(p_dev, s), capi.solve(p_host, s)
    assert a.termination_status == b.termination_status == 1
    assert np.array_equal(a.primal(), b.primal())  # same arrays in, deterministic solver
    # getter into a device buffer
    x_dev = torch.zeros(lp.n, dtype=torch.float64, device=dev)
    assert L.cuOptGetPrimalSolution(a.h, ptr(x_dev)) == 0
    assert np.array_equal(x_dev.cpu().numpy(), a.primal())

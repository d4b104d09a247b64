"""The reference's OWN C-API test code run against this repo's libcuopt.so.

oracle/_ref/libref_c_api_test.so is `cpp/tests/linear_programming/c_api_tests/c_api_test.c` of the reference, compiled
unchanged from where it lies (oracle/Makefile, target `ref`) against THIS repo's include/ and linked to THIS repo's
library — so the file compiling at all checks the header, and each function below is one TEST of the reference's
c_api_tests.cpp with the same expectation (cited).  The library is built where /root/reference exists and travels to
the GPU box as a prebuilt file; without it the tests skip.
Not reproduced: `burglar` (a MIP: this LP-only build answers with CUOPT_VALIDATION_ERROR, asserted below) and the
time-limit fixture (its instances square41 / enlight_hard are not in the tree)."""
import ctypes as C
import os

import pytest

from conftest import mps_path

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "..", "oracle", "_ref", "libref_c_api_test.so")
INF = float("inf")
SUCCESS, INVALID_ARGUMENT, MPS_FILE_ERROR, VALIDATION_ERROR = 0, 1, 2, 4
OPTIMAL, INFEASIBLE, ITERATION_LIMIT = 1, 2, 4


@pytest.fixture(scope="module")
def ref():
    if not os.path.exists(LIB):
        pytest.skip("oracle/_ref/libref_c_api_test.so not built (needs /root/reference: make -C oracle ref)")
    L = C.CDLL(LIB)
    L.solve_mps_file.argtypes = [C.c_char_p, C.c_double, C.c_double, C.POINTER(C.c_int), C.POINTER(C.c_double), C.c_int]
    L.test_ranged_problem.argtypes = [C.POINTER(C.c_int), C.POINTER(C.c_double)]
    return L


# ---- no GPU needed -------------------------------------------------------------------------------------------
def test_int_and_float_size(ref):            # TEST(c_api, int_size) / TEST(c_api, float_size)
    assert ref.test_int_size() == 4 and ref.test_float_size() == 8


def test_missing_file(ref):                  # TEST(c_api, test_missing_file)
    assert ref.test_missing_file() == MPS_FILE_ERROR


def test_bad_parameter_name(ref):            # TEST(c_api, bad_parameter_name)
    assert ref.test_bad_parameter_name() == INVALID_ARGUMENT


def test_burglar_is_a_mip_and_is_refused(ref):   # TEST(c_api, burglar) expects SUCCESS from the reference's MIP solver;
    assert ref.burglar_problem() == VALIDATION_ERROR  # this LP-only build refuses before touching the GPU


# ---- on the GPU ----------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_afiro(ref):                         # TEST(c_api, afiro): default method of solve_mps_file = DUAL_SIMPLEX
    status = C.c_int(-1)
    rc = ref.solve_mps_file(mps_path("linear_programming/afiro_original.mps").encode(), 60.0, INF, C.byref(status),
                            None, 2)
    assert rc == SUCCESS and status.value == OPTIMAL


@pytest.mark.gpu
def test_iteration_limit(ref):               # TEST(c_api, iteration_limit)
    status = C.c_int(-1)
    rc = ref.solve_mps_file(mps_path("linear_programming/afiro_original.mps").encode(), 60.0, 1.0, C.byref(status),
                            None, 2)
    assert rc == SUCCESS and status.value == ITERATION_LIMIT


@pytest.mark.gpu
def test_infeasible_problem(ref):            # TEST(c_api, test_infeasible_problem): checks the status itself
    assert ref.test_infeasible_problem() == SUCCESS


@pytest.mark.gpu
def test_ranged_problem(ref):                # TEST(c_api, test_ranged_problem)
    status, objective = C.c_int(-1), C.c_double(0.0)
    assert ref.test_ranged_problem(C.byref(status), C.byref(objective)) == SUCCESS
    assert status.value == OPTIMAL
    assert objective.value == pytest.approx(32.0, abs=1e-3)



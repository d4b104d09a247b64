"""MPS ingest parity: cuOptReadProblem (our reader, cuopt_b200/csrc/mps_reader.cpp) against dumps produced by the
REFERENCE parser (cpp/libmps_parser, run by scripts/gen_golden.py) for every MPS file of the reference's datasets/,
in free and fixed mode.  Mirrors cpp/libmps_parser/tests/mps_parser_test.cpp (good files parse to the same arrays,
bad files are rejected)."""
import hashlib

import numpy as np
import pytest

from conftest import load_golden, mps_path
from cuopt_b200 import capi

GOLD = load_golden("parser_golden.json")


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def _inf_list(v):
    return np.array([np.inf if x == "inf" else -np.inf if x == "-inf" else x for x in v], dtype=float)


@pytest.mark.parametrize("mode", ["free", "fixed"])
@pytest.mark.parametrize("rel", sorted(GOLD))
def test_reader_matches_reference_parser(rel, mode, tmp_path):
    want = GOLD[rel][mode]
    path = mps_path(rel, str(tmp_path))
    if not want["ok"]:
        with pytest.raises(capi.CuOptError) as e:
            capi.Problem.read(path, fixed_format=(mode == "fixed"))
        assert e.value.code == capi.CUOPT_MPS_PARSE_ERROR
        return
    p = capi.Problem.read(path, fixed_format=(mode == "fixed"))
    assert (p.num_constraints, p.num_variables, p.num_nonzeros) == (want["m"], want["n"], want["nnz"])
    assert (p.objective_sense == capi.CUOPT_MAXIMIZE) == want["maximize"]
    assert p.objective_offset == want["objective_offset"]
    off, idx, val = p.constraint_matrix()
    got = dict(offsets=off, indices=idx, values=val, rhs=p.rhs(), c=p.objective_coefficients(),
               var_lb=p.variable_lower_bounds(), var_ub=p.variable_upper_bounds(),
               con_lb=p.constraint_lower_bounds(), con_ub=p.constraint_upper_bounds())
    for k, d in want["digests"].items():
        assert digest(got[k]) == d, f"{rel} [{mode}] array {k} differs from the reference parser"
    assert p.variable_types().decode() == want["var_types"]
    assert p.is_mip == ("I" in want["var_types"])
    if "arrays" in want:  # small files: element-wise too (readable failures)
        np.testing.assert_array_equal(got["con_lb"], _inf_list(want["arrays"]["con_lb"]))
        np.testing.assert_array_equal(got["var_ub"], _inf_list(want["arrays"]["var_ub"]))


def test_read_problem_free_is_default():
    p = capi.Problem.read(mps_path("linear_programming/afiro_original.mps"))
    assert (p.num_constraints, p.num_variables, p.num_nonzeros) == (27, 32, 83)
    # like the reference data model, an MPS problem is carried in ranged form: no sense characters
    assert not np.isnan(p.constraint_lower_bounds()).any()


def test_missing_file_and_bad_file_codes(tmp_path):
    # c_api_tests: missing file -> CUOPT_MPS_FILE_ERROR, malformed -> CUOPT_MPS_PARSE_ERROR (cuopt_c.cpp:71-79)
    with pytest.raises(capi.CuOptError) as e:
        capi.Problem.read(str(tmp_path / "does_not_exist.mps"))
    assert e.value.code == capi.CUOPT_MPS_FILE_ERROR
    bad = tmp_path / "bad.mps"
    bad.write_text("NAME x\nROWS\n N obj\nFOO\nENDATA\n")
    with pytest.raises(capi.CuOptError) as e:
        capi.Problem.read(str(bad))
    assert e.value.code == capi.CUOPT_MPS_PARSE_ERROR

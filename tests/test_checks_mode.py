"""The reference driver's self-test `--checks` (laghos.cpp:904-926, 1403-1474) behind the C ABI: the table is the
reference's (it_norms, laghos.cpp:1441-1463) -- its transcription in host/checks.hpp is compared with
tests/golden/checks_table.json -- and the step logic (two-sided relative error below eps, two hits per run) is driven
here by CPU oracle runs at the reference's own tolerance 1e-13."""
import ctypes as C
import json
import os

import pytest

import pyoracle
from laghos_b200 import load_library

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TABLE = json.load(open(os.path.join(ROOT, "tests", "golden", "checks_table.json")))["parallel"]


def test_table_transcription(built):
    lib = load_library()
    for dim in (2, 3):
        for problem in range(8):
            for k in range(2):
                it, nrm = C.c_int32(), C.c_double()
                assert lib.lagb_checks_entry(dim, problem, k, C.byref(it), C.byref(nrm)) == 0
                assert [it.value, nrm.value] == TABLE[str(dim)][str(problem)][k]
    assert lib.lagb_checks_entry(4, 0, 0, C.byref(it), C.byref(nrm)) != 0


def test_step_logic(built):
    lib = load_library()
    chk = C.c_int32(0)
    it, ref = TABLE["3"]["1"][0]
    assert lib.lagb_checks_step(3, 1, it, ref, 1e-13, C.byref(chk)) == 0 and chk.value == 1
    assert lib.lagb_checks_step(3, 1, it + 1, ref, 1e-13, C.byref(chk)) == 0 and chk.value == 1    # no entry at this step
    assert lib.lagb_checks_step(3, 1, it, ref * (1 + 5e-14), 1e-13, C.byref(chk)) == 0 and chk.value == 2
    assert lib.lagb_checks_step(3, 1, it, ref * (1 + 3e-13), 1e-13, C.byref(chk)) != 0
    assert lib.lagb_last_error().decode() == f"P1, #{it}" and chk.value == 3      # the reference counts before it verifies
    assert lib.lagb_checks_step(3, 1, it, 0.0, 1e-13, C.byref(chk)) != 0
    assert "near zero" in lib.lagb_last_error().decode()


@pytest.mark.parametrize("dim,problem", [(2, 1), (2, 3), (3, 3)])
def test_oracle_run_passes_the_reference_check(built, dim, problem):
    lib = load_library()
    mesh = "square01_quad" if dim == 2 else "cube01_hex"
    r = pyoracle.run(mesh=mesh, rs=0, problem=problem, ok=2, ot=1, t_final=0.6, cfl=0.5, cg_tol=1e-14, hist_cap=4096)
    chk = C.c_int32(0)
    for ti, nrm in r["hist"]:
        assert lib.lagb_checks_step(dim, problem, ti, nrm, 1e-13, C.byref(chk)) == 0, lib.lagb_last_error()
    assert chk.value == 2                                    # laghos.cpp:926

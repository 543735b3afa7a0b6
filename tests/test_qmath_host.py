"""Host check of the device point-physics math (qmath.cuh: trig-free cubic root, restructured symmetric
eigen-solver and smallest singular value) against the oracle's restatement of MFEM's routines
(oracle/smallmat.hpp, pinned on the reference's golden values through tests/test_oracle_golden.py).
The functions are __host__ __device__; the device-only hardware seeds (rcp/rsqrt.approx) are covered by
the GPU parity tests (tests/test_gpu_operators.py::test_qupdate_and_force)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_qmath_host_vs_oracle(tmp_path):
    exe = tmp_path / "qmath_check"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I/usr/local/cuda/include", "-w",
                           os.path.join(ROOT, "tests", "cpp", "qmath_host_check.cpp"), "-o", str(exe)])
    out = subprocess.check_output([str(exe)], text=True)
    vals = {l.split()[0]: l.split()[1:] for l in out.strip().splitlines()}
    assert float(vals["cos_acos_third"][0]) < 4e-16          # <= 2 ulp of cos(acos(x)/3) on [-0.9, 1]
    assert float(vals["min_eig3_value"][0]) < 1e-13           # relative to the matrix norm, degenerate spectra included
    assert float(vals["min_eig3_vector"][0]) < 1e-9 and int(vals["min_eig3_vector"][1]) > 100000
    assert float(vals["min_sv3"][0]) < 1e-12
    assert [float(v) for v in vals["zero"]] == [0.0, 1.0, 0.0, 0.0]

"""End-to-end parity on the GPU: the restated driver loop (C++ shim over the C ABI) against
the reference's known answers and against the oracle.

* --checks table (laghos.cpp:1441-1463): rel. tolerance 1e-13 in the reference for CPU/GPU/any
  rank count at -cgt 1e-14.  The B200 path sums in a different order (atomics, batched PCG), so the
  gate here is 1e-11; the north_star bar is 1e-9 on |e|.
* Q3Q2 3D Sedov / Taylor-Green (BASELINE configs 2/3 at reduced -rs): |e| vs the oracle at equal
  step index with identical accepted/rejected step sequences, rel. 1e-9.
"""
import json
import os

import pytest

import pyoracle

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
TABLE = json.load(open(os.path.join(HERE, "golden", "checks_table.json")))


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("problem", range(8))
def test_checks_table_gpu(built, dim, problem, variant):
    from laghos_b200.api import run
    if variant == 1 and problem not in (0, 1):
        pytest.skip("generic variant: problems 0 and 1 only (time)")
    mesh = "square01_quad" if dim == 2 else "cube01_hex"
    r = run(mesh=mesh, rs=0, problem=problem, ok=2, ot=1, t_final=0.6, cfl=0.5, cg_tol=1e-14,
            kernel_variant=variant, hist_cap=4096)
    hist = dict(r["hist"])
    for it, ref in TABLE["parallel"][str(dim)][str(problem)]:
        assert it in hist
        rel = abs(hist[it] - ref) / abs(ref)
        assert rel < 1e-11, (dim, problem, it, hist[it], ref, rel)


@pytest.mark.parametrize("batched", [True, False])
@pytest.mark.parametrize("cfg", [
    dict(mesh="cube01_hex", rs=2, problem=1, ok=3, ot=2, max_tsteps=12),   # BASELINE config 2 at rs 2
    dict(mesh="cube01_hex", rs=2, problem=0, ok=3, ot=2, max_tsteps=12),   # BASELINE config 3 at rs 2
    dict(mesh="box01_hex", rs=1, problem=3, ok=2, ot=1, max_tsteps=12),    # BASELINE config 5, ok 2
    dict(mesh="square01_quad", rs=3, problem=0, ok=2, ot=1, max_tsteps=20),  # BASELINE config 1
    dict(mesh="box01_hex", rs=1, problem=3, ok=3, ot=2, max_tsteps=6),     # BASELINE config 5, ok 3
    dict(mesh="box01_hex", rs=0, problem=3, ok=4, ot=3, max_tsteps=6),     # BASELINE config 5, ok 4
    # BASELINE config 5, ok 5 (no reference kernel).  The unpreconditioned L2 CG on the order-4 Bernstein mass
    # matrix needs ~600 iterations at -cgt 1e-12: with the default -cgm 300 it stops unconverged and even the
    # oracle with 1 vs 8 threads differs by 3e-5 on |e|.  Run this case with -cgm 2000; two round-off paths
    # then agree (oracle 1 vs 8 threads: 1e-14): tolerance 1e-8 on |e|.
    dict(mesh="box01_hex", rs=0, problem=3, ok=5, ot=4, max_tsteps=4, tol=1e-8, cg_max_iter=2000),
], ids=["sedov-q3q2", "tg-q3q2", "triple-q2q1", "tg2d-q2q1", "triple-q3q2", "triple-q4q3", "triple-q5q4"])
def test_vs_oracle_e_norm(built, cfg, batched):
    from laghos_b200.api import run
    kw = dict(cfg, t_final=10.0, cg_tol=1e-12)
    tol = kw.pop("tol", 1e-9)
    ro = pyoracle.run(**kw, nthreads=8)
    rg = run(**kw, batched_pcg=batched, hist_cap=4096)
    assert rg["steps"] == ro["steps"] and rg["ti_last"] == ro["ti_last"]
    assert len(rg["hist"]) == len(ro["hist"])
    for (ti_g, e_g), (ti_o, e_o) in zip(rg["hist"], ro["hist"]):
        assert ti_g == ti_o
        assert abs(e_g - e_o) <= tol * abs(e_o), (ti_g, e_g, e_o)
    assert abs(rg["dt"] - ro["dt"]) <= tol * ro["dt"]
    # total energy IE + KE before / after (the reference's "Energy diff" line, laghos.cpp:956-962)
    assert abs(rg["energy_init"] - ro["energy_init"]) <= 1e-12 * abs(ro["energy_init"])
    assert abs(rg["energy_final"] - ro["energy_final"]) <= max(tol, 1e-9) * abs(ro["energy_final"])


def test_readme_run2_gpu(built):
    """README.md:216,228: -p 0 -m cube01_hex -rs 1 -tf 0.75 -pa -> 1041 steps, dt 0.000121, |e| 3.3909635545e+03."""
    from laghos_b200.api import run
    g = TABLE["readme"]["run2"]
    r = run(**g["args"])
    assert r["ti_last"] == g["step"]
    assert f"{r['dt']:.6f}" == g["dt"]
    assert f"{r['e_norm']:.10e}" == g["e_norm"]


def test_readme_run8_rk2avg_gpu(built):
    """README.md:222,234: -p 4 -m square_gresho -rs 3 -ok 3 -ot 2 -tf 0.62831853 -s 7 -pa (RK2Avg, Q3Q2 2D):
    776 steps, dt 0.000045, |e| 4.0982431726e+02."""
    from laghos_b200.api import run
    g = TABLE["readme"]["run8"]
    r = run(**g["args"])
    assert r["ti_last"] == g["step"]
    assert f"{r['dt']:.6f}" == g["dt"]
    assert f"{r['e_norm']:.10e}" == g["e_norm"]


def test_e2e_host_state_matches_resident(built):
    from laghos_b200.api import run
    # -cgt 1e-13: at the default 1e-8 the scatter's atomic summation order moves |e| by ~1e-12 run to run
    kw = dict(mesh="cube01_hex", rs=1, problem=1, ok=3, ot=2, max_tsteps=5, t_final=10.0, cg_tol=1e-13)
    a = run(**kw)
    b = run(**kw, e2e_host_state=True)
    assert a["steps"] == b["steps"]
    assert abs(a["e_norm"] - b["e_norm"]) <= 1e-11 * abs(a["e_norm"])
    assert b["h2d_bytes_per_step"] > 0 and b["d2h_bytes_per_step"] > 0


@pytest.mark.parametrize("ode", [1, 2, 3, 6, 7], ids=["euler", "rk2", "rk3ssp", "rk6", "rk2avg"])
def test_ode_solver_types_vs_oracle(built, ode):
    """Every -s value of the reference (laghos.cpp:519-534; 4 = RK4 is covered above, 6 = MFEM's 8-stage RK6):
    |e| after every step against the oracle's restatement of the same integrator."""
    from laghos_b200.api import run
    kw = dict(mesh="cube01_hex", rs=1, problem=1, ok=2, ot=1, max_tsteps=5, t_final=10.0, cg_tol=1e-12,
              ode_solver_type=ode)
    ro = pyoracle.run(**kw, nthreads=4)
    rg = run(**kw, hist_cap=64)
    assert rg["steps"] == ro["steps"] and len(rg["hist"]) == len(ro["hist"])
    for (ti_g, e_g), (ti_o, e_o) in zip(rg["hist"], ro["hist"]):
        assert ti_g == ti_o and abs(e_g - e_o) <= 1e-9 * abs(e_o), (ode, ti_g, e_g, e_o)

"""Pin the CPU oracle on the reference's own known answers.

* the --checks table, reference laghos.cpp:1441-1463 (rel. tolerance 1e-13, -rs 0 -ok 2 -ot 1 -s 4
  -cfl 0.5 -tf 0.6 -cgt 1e-14; makefile:199), serial-driver variants serial/laghos.cpp:803-869;
* the end-of-run table README.md:225-235 / makefile:271-278 (default -cgt 1e-8).
The same table is committed as tests/golden/checks_table.json.
"""
import json
import os

import pytest

import pyoracle

HERE = os.path.dirname(os.path.abspath(__file__))
TABLE = json.load(open(os.path.join(HERE, "golden", "checks_table.json")))


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("problem", range(8))
def test_checks_table(built, dim, problem):
    mesh = "square01_quad" if dim == 2 else "cube01_hex"
    entries = TABLE["parallel"][str(dim)][str(problem)]
    r = pyoracle.run(mesh=mesh, rs=0, problem=problem, ok=2, ot=1, t_final=0.6, cfl=0.5, cg_tol=1e-14)
    hist = dict(r["hist"])
    for it, ref in entries:
        assert it in hist, f"iteration {it} not reached"
        rel = abs(hist[it] - ref) / abs(ref)
        assert rel < 1e-13, (dim, problem, it, hist[it], ref, rel)


def test_checks_serial_driver_3d_sedov(built):
    # serial/laghos.cpp:101,319: blast energy 0.25, not divided by 2^dim
    ent = TABLE["serial"]["3"]["1"]
    r = pyoracle.run(mesh="cube01_hex", rs=0, problem=1, blast_scale=0.25, t_final=0.6, cg_tol=1e-14)
    hist = dict(r["hist"])
    for it, ref in ent:
        assert abs(hist[it] - ref) / abs(ref) < 1e-13


README_RUNS = TABLE["readme"]


def _check_readme(run):
    kw = dict(run["args"])
    r = pyoracle.run(**kw)
    assert r["ti_last"] == run["step"], (r["ti_last"], run["step"])
    assert f"{r['dt']:.6f}" == run["dt"], (r["dt"], run["dt"])
    assert f"{r['e_norm']:.10e}" == run["e_norm"], (r["e_norm"], run["e_norm"])


@pytest.mark.parametrize("name", ["run2"])
def test_readme_end_of_run_fast(built, name):
    _check_readme(README_RUNS[name])


@pytest.mark.slow
@pytest.mark.parametrize("name", ["run1", "run3", "run4", "run6", "run7", "run8", "run9"])
def test_readme_end_of_run_slow(built, name):
    _check_readme(README_RUNS[name])

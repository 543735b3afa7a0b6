"""GPU parity at benchmark scale: the small-mesh tests never reach the 64-bit index
paths ((size_t)NE*NQ*9 = 5.1e8 at -rs 5), grids of 32768+ CTAs or the PCG partial-buffer sizing.

* operator level at cube01_hex -rs 4 -ok 3 -ot 2 (32768 elements) against the CPU oracle run here;
* |e| after every one of the first RK4 steps at -rs 4 and at the BASELINE size -rs 5 (262144 elements)
  against tests/golden/bench_enorm.json (CPU oracle, tools/make_bench_golden.py), north_star bar 1e-9.
"""
import json
import os

import numpy as np
import pytest

import pyoracle

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "bench_enorm.json")))


def relerr(a, b):
    a = np.asarray(a); b = np.asarray(b)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def test_operators_rs4(built):
    from laghos_b200.api import Problem, Context
    mesh, rs, problem, ok, ot = "cube01_hex", 4, 1, 3, 2
    nthr = os.cpu_count() or 1
    P = Problem(mesh, rs, problem, ok, ot)
    O = pyoracle.Oracle(mesh, rs, problem, ok, ot, nthreads=nthr)
    c = Context(P)
    try:
        rng = np.random.default_rng(4)
        nv, nd, nl = P.h1_vsize, P.ndofs_h1, P.ndofs_l2
        # Q: IC and a perturbed state (velocity everywhere: the general eigen-solver paths)
        S = P.S0.copy()
        n1 = int(P.info.nelem[0])
        S[:nv] += 0.1 * (0.5 / (n1 * 3)) * rng.uniform(-1, 1, nv)
        S[nv:2 * nv] = rng.uniform(-1, 1, nv)
        S[2 * nv:] = rng.uniform(0.5, 1.5, nl)
        for state in (P.S0.copy(), S):
            dt_ref = O.qupdate(state)
            dt = c.qupdate(c.dev(state))
            assert abs(dt - dt_ref) <= 1e-13 * abs(dt_ref), (dt, dt_ref)
            assert relerr(c.qdata(0).cpu().numpy(), O.qdata(0)) < 1e-11
        # F, F^T on the quadrature data of the perturbed state
        e = rng.uniform(0.5, 1.5, nl)
        v = rng.uniform(-1, 1, nv)
        assert relerr(c.force_mult(c.dev(e)).cpu().numpy(), O.force_mult(e)) < 1e-11
        assert relerr(c.force_mult_transpose(c.dev(v)).cpu().numpy(), O.force_mult_transpose(v)) < 1e-11
        # M: batched apply (the PCG's operator) and the single-component apply with essential dofs
        y = c.vmass_mult_all(c.dev(v)).cpu().numpy()
        for comp in range(3):
            ref = O.vmass_mult(v[comp * nd:(comp + 1) * nd].copy(), -1)
            assert relerr(y[comp * nd:(comp + 1) * nd], ref) < 1e-12
        assert relerr(c.vmass_mult(c.dev(v[:nd].copy()), 0).cpu().numpy(), O.vmass_mult(v[:nd].copy(), 0)) < 1e-12
        # batched PCG against the oracle's scalar solves
        xa, its = c.pcg_vmass_all(c.dev(v))
        xa = xa.cpu().numpy()
        for comp in range(3):
            xr, itr = O.pcg_vmass(comp, v[comp * nd:(comp + 1) * nd].copy())
            assert abs(its[comp] - itr) <= 1
            assert relerr(xa[comp * nd:(comp + 1) * nd], xr) < 1e-7
    finally:
        c.close()


@pytest.mark.parametrize("rs,steps", [(4, 4), (5, 3)], ids=["rs4", "rs5-baseline-size"])
def test_first_steps_match_oracle_golden(built, rs, steps):
    from laghos_b200.api import run
    gold = GOLD[f"sedov_rs{rs}"]
    r = run(mesh="cube01_hex", rs=rs, problem=1, ok=3, ot=2, t_final=1e9, cg_tol=gold["cg_tol"],
            max_tsteps=steps - 1, hist_cap=64)
    assert r["steps"] == steps
    hist = dict(r["hist"])
    for ti in range(1, steps + 1):
        ref = gold["e_norm_after_step"][str(ti)]
        assert abs(hist[ti] - ref) <= 1e-9 * abs(ref), (rs, ti, hist[ti], ref)

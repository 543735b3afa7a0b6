"""GPU tests of the coloured brick mass apply (device/mass3d_brick.cuh, host/batch_plan.hpp):
every launch variant, structured and unstructured batching, against the oracle and against the
legacy atomic-scatter kernel; bitwise run-to-run reproducibility (the reference's E^t is
deterministic, MFEM ElementRestriction::MultTranspose behind laghos_assembly.cpp:117-121)."""
import numpy as np
import pytest

import pyoracle

pytestmark = pytest.mark.gpu


def relerr(a, b):
    a = np.asarray(a); b = np.asarray(b)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


@pytest.mark.parametrize("mesh,rs,ok", [("cube01_hex", 2, 3), ("box01_hex", 1, 3), ("cube01_hex", 2, 2), ("cube01_hex", 1, 4),
                                        ("cube01_hex", 2, 1), ("cube01_hex", 0, 5)])
def test_brick_variants_vs_oracle(built, mesh, rs, ok):
    from laghos_b200.api import Problem, Context
    P = Problem(mesh, rs, 1, ok, ok - 1)
    O = pyoracle.Oracle(mesh, rs, 1, ok, ok - 1)
    rng = np.random.default_rng(21)
    x = rng.uniform(-1, 1, P.h1_vsize)
    ref = np.concatenate([O.vmass_mult(x[c * P.ndofs_h1:(c + 1) * P.ndofs_h1], -1) for c in range(3)])
    for hint in (True, False):
        c = Context(P, grid_hint=hint)
        for path, shape in ((4, 0), (4, 1), (4, 2), (3, 0), (3, 1), (3, 2), (2, 0)):   # dataflow kernel / brick v2 with the three brick shapes, brick v1
            c.tune(6, path)
            c.tune(7, shape)
            for var in range(5 if path >= 3 else 3):
                c.tune(4, var)
                for pdl_off in (0, 1):
                    c.tune(5, pdl_off)
                    y = c.empty(P.h1_vsize)
                    y.fill_(float("nan"))          # the schedule must not depend on the output's previous contents
                    c.vmass_mult_all(c.dev(x), y)
                    y2 = c.vmass_mult_all(c.dev(x)).cpu().numpy()
                    y = y.cpu().numpy()
                    assert relerr(y, ref) < 1e-12, (hint, path, shape, var, pdl_off)
                    assert np.array_equal(y, y2), "brick apply must be bitwise reproducible"
                y1 = c.vmass_mult(c.dev(x[:P.ndofs_h1]), 0).cpu().numpy()
                r1 = ref[:P.ndofs_h1].copy(); r1[P.ess(0)] = 0.0
                assert relerr(y1, r1) < 1e-12, (hint, path, shape, var)
        c.tune(4, 0); c.tune(5, 0); c.tune(7, 0)
        c.tune(6, 1)                           # legacy atomic path still agrees
        assert relerr(c.vmass_mult_all(c.dev(x)).cpu().numpy(), ref) < 1e-12
        c.close()


@pytest.mark.parametrize("mesh,rs,ok", [("cube01_hex", 2, 3), ("cube01_hex", 1, 2)])
def test_brick_pcg_matches_legacy_and_oracle(built, mesh, rs, ok):
    """The fused-direction PCG produces the iterates of the plain one (MFEM CGSolver semantics)."""
    from laghos_b200.api import Problem, Context
    P = Problem(mesh, rs, 1, ok, ok - 1)
    O = pyoracle.Oracle(mesh, rs, 1, ok, ok - 1, cg_tol=1e-14)
    rng = np.random.default_rng(22)
    b = rng.uniform(-1, 1, P.h1_vsize)
    refs = [O.pcg_vmass(comp, b[comp * P.ndofs_h1:(comp + 1) * P.ndofs_h1].copy()) for comp in range(3)]
    c = Context(P)
    sols = {}
    for legacy in (4, 3, 2, 1):                # dataflow kernel (plain PCG), brick v2 / v1 (fused PCG), legacy atomic scatter
        c.tune(6, legacy)
        xa, its = c.pcg_vmass_all(c.dev(b), rel_tol=1e-14)
        xa2, its2 = c.pcg_vmass_all(c.dev(b), rel_tol=1e-14)
        sols[legacy] = xa.cpu().numpy()
        if legacy >= 2:
            assert np.array_equal(sols[legacy], xa2.cpu().numpy()) and list(its) == list(its2), "brick PCG must be reproducible"
        for comp in range(3):
            xr, itr = refs[comp]
            assert abs(its[comp] - itr) <= 1
            assert relerr(sols[legacy][comp * P.ndofs_h1:(comp + 1) * P.ndofs_h1], xr) < 1e-11
            assert np.all(sols[legacy][comp * P.ndofs_h1:(comp + 1) * P.ndofs_h1][P.ess(comp)] == 0.0)
        x0, it0 = c.pcg_vmass(1, c.dev(b[P.ndofs_h1:2 * P.ndofs_h1]), rel_tol=1e-14)
        assert relerr(x0.cpu().numpy(), refs[1][0]) < 1e-11
    assert relerr(sols[3], sols[1]) < 1e-12 and relerr(sols[2], sols[1]) < 1e-12 and relerr(sols[4], sols[1]) < 1e-12
    # the split vector kernels (x updated in the direction kernel) reproduce the first version bit for bit
    c.tune(6, 4)
    c.tune(8, 1)
    xo, ito = c.pcg_vmass_all(c.dev(b), rel_tol=1e-14)
    assert np.array_equal(xo.cpu().numpy(), sols[4]) and list(ito) == list(its)
    c.tune(8, 0)
    c.tune(9, 1)                               # plain PCG on the multi-launch brick kernel
    c.tune(6, 3)
    xp, itp = c.pcg_vmass_all(c.dev(b), rel_tol=1e-14)
    assert relerr(xp.cpu().numpy(), sols[1]) < 1e-12
    c.close()

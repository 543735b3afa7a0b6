"""End-of-run velocity error norms (reference laghos.cpp:970-982, problems 0 and 4; host/error_norms.hpp) against an
independent numpy restatement (numpy's Gauss-Legendre, barycentric Lagrange on Gauss-Lobatto nodes built from the
roots of P'_n) and a closed-form anchor: with v_h = 0 on the unit square, L_2 = |v0|_{L2} = sqrt(1/2) for the 2D
Taylor-Green field."""
import numpy as np
import pytest
from numpy.polynomial import legendre as L

from laghos_b200.api import Problem
from test_tables_independent import gauss01, gll01


def lagrange(nodes, x):
    n = len(nodes)
    B = np.ones((len(x), n))
    G = np.zeros((len(x), n))
    for j in range(n):
        for m in range(n):
            if m != j:
                B[:, j] *= (x - nodes[m]) / (nodes[j] - nodes[m])
        for k in range(n):
            if k == j:
                continue
            t = np.full(len(x), 1.0 / (nodes[j] - nodes[k]))
            for m in range(n):
                if m not in (j, k):
                    t *= (x - nodes[m]) / (nodes[j] - nodes[m])
            G[:, j] += t
    return B, G


def v0(problem, X):
    x, y = X[0], X[1]
    v = np.zeros_like(X)
    if problem == 0:
        v[0] = np.sin(np.pi * x) * np.cos(np.pi * y)
        v[1] = -np.cos(np.pi * x) * np.sin(np.pi * y)
        if X.shape[0] == 3:
            v[0] *= np.cos(np.pi * X[2])
            v[1] *= np.cos(np.pi * X[2])
    elif problem == 4:                                    # Gresho vortex, laghos.cpp:1180-1195
        r = np.sqrt(x * x + y * y)
        with np.errstate(divide="ignore", invalid="ignore"):
            v[0] = np.where(r < 0.2, 5 * y, np.where(r < 0.4, 2 * y / r - 5 * y, 0.0))
            v[1] = np.where(r < 0.2, -5 * x, np.where(r < 0.4, -2 * x / r + 5 * x, 0.0))
    return v


def numpy_errors(P, S, problem):
    dim, D, ok = P.dim, P.D1D, P.D1D - 1
    gx, gw = gauss01(ok + 2)
    B, G = lagrange(gll01(D), gx)
    nd = P.ndofs_h1
    X = S[:dim * nd].reshape(dim, nd)[:, P.h1_map].reshape((dim, P.NE) + (D,) * dim)
    V = S[dim * nd:2 * dim * nd].reshape(dim, nd)[:, P.h1_map].reshape((dim, P.NE) + (D,) * dim)
    # local index = kx + D (ky + D kz): the LAST array axis is x
    if dim == 2:
        ev = lambda T, Ay, Ax: np.einsum("ceyx,qy,px->ceqp", T, Ay, Ax)
        xq, vq = ev(X, B, B), ev(V, B, B)
        J = np.stack([ev(X, B, G), ev(X, G, B)], axis=1)            # [c, d, e, qy, qx]
        det = J[0, 0] * J[1, 1] - J[0, 1] * J[1, 0]
        w = np.einsum("q,p->qp", gw, gw)
    else:
        ev = lambda T, Az, Ay, Ax: np.einsum("cezyx,rz,qy,px->cerqp", T, Az, Ay, Ax)
        xq, vq = ev(X, B, B, B), ev(V, B, B, B)
        J = np.stack([ev(X, B, B, G), ev(X, B, G, B), ev(X, G, B, B)], axis=1)
        det = np.linalg.det(np.moveaxis(J, (0, 1), (-2, -1)))
        w = np.einsum("r,q,p->rqp", gw, gw, gw)
    err = np.sqrt(((vq - v0(problem, xq)) ** 2).sum(axis=0))
    return err.max(), (w * det * err).sum(), np.sqrt((w * det * err ** 2).sum())


@pytest.mark.parametrize("mesh,rs,problem,ok", [("square01_quad", 2, 0, 2), ("square01_quad", 1, 0, 4),
                                                ("cube01_hex", 1, 0, 3), ("square_gresho", 1, 4, 3),
                                                ("cube01_hex", 2, 0, 2)])      # 512 elements: the threaded element loop
def test_against_numpy(built, mesh, rs, problem, ok):
    P = Problem(mesh=mesh, rs=rs, problem=problem, ok=ok, ot=ok - 1)
    rng = np.random.default_rng(11)
    S = np.array(P.S0)
    nv = P.dim * P.ndofs_h1
    h = 1.0 / (2 ** rs * 2 * ok)
    S[:nv] += 0.05 * h * rng.standard_normal(nv)          # a deformed (still valid) mesh
    S[nv:2 * nv] += 0.1 * rng.standard_normal(nv)
    got = P.velocity_error(S)
    ref = numpy_errors(P, S, problem)
    assert np.allclose(got, ref, rtol=1e-12, atol=1e-14), (got, ref)


def test_closed_form_and_convergence(built):
    # v_h = 0 on the initial mesh: L_2 = |v0|_{L2(unit square)} = sqrt(1/2), L_inf -> 1 (|v0| = 1 on the edge midpoints)
    P = Problem(mesh="square01_quad", rs=2, problem=0, ok=3, ot=2)
    S = np.array(P.S0)
    nv = P.dim * P.ndofs_h1
    S[nv:2 * nv] = 0.0
    linf, l1, l2 = P.velocity_error(S)
    assert abs(l2 - np.sqrt(0.5)) < 1e-9 and 0.95 < linf <= 1.0 + 1e-12 and 0 < l1 < l2
    # at t = 0 the error is the interpolation error of v0: it falls with the order (the boundary condition sets
    # v.n = 0 exactly where v0.n = 0, so no O(1) boundary term)
    errs = []
    for ok in (2, 3, 4):
        Q = Problem(mesh="square01_quad", rs=2, problem=0, ok=ok, ot=ok - 1)
        errs.append(Q.velocity_error(np.array(Q.S0))[2])
    assert errs[0] > 5 * errs[1] > 25 * errs[2] > 0
    # problems whose v0 is zero: the norms of v itself
    R = Problem(mesh="cube01_hex", rs=0, problem=1, ok=2, ot=1)
    S = np.array(R.S0)
    S[3 * R.ndofs_h1:4 * R.ndofs_h1] = 2.0                 # v_y = 2 everywhere on the unit cube
    assert np.allclose(R.velocity_error(S), (2.0, 2.0, 2.0), rtol=1e-13)


@pytest.mark.parametrize("pgrid", [(2, 1, 1), (2, 2, 1), (2, 2, 2)])
def test_rank_blocks_add_up(built, pgrid):
    """partitioned runs reduce the per-rank values (max / sum / sum of squares): the blocks of the initial state must
    reproduce the one-rank norms -- velocity errors (problem 0) and the Sedov density error (problem 1)"""
    import ctypes as C
    nr = pgrid[0] * pgrid[1] * pgrid[2]
    cfg = dict(mesh="cube01_hex", rs=1, ok=2, ot=1)
    G = Problem(problem=0, **cfg)
    S = np.array(G.S0)
    S[3 * G.ndofs_h1:6 * G.ndofs_h1] *= 0.9                 # v = 0.9 v0: a non-trivial error field, the same on every rank
    ref = G.velocity_error(S)
    mx, l1, l2sq = 0.0, 0.0, 0.0
    for rank in range(nr):
        P = Problem(problem=0, rank=rank, pgrid=pgrid, **cfg)
        Sl = np.array(P.S0)
        Sl[3 * P.ndofs_h1:6 * P.ndofs_h1] *= 0.9
        out = (C.c_double * 4)()
        assert P.lib.lagb_problem_velocity_error(P.h, Sl.ctypes.data_as(C.c_void_p), out) == 0
        mx, l1, l2sq = max(mx, out[0]), l1 + out[1], l2sq + out[3]
    assert np.allclose((mx, l1, np.sqrt(l2sq)), ref, rtol=1e-12)
    G = Problem(problem=1, **cfg)
    ref = G.sedov_density_error(np.array(G.S0), np.array(G.rho0_gf), 0.2)
    tot = 0.0
    for rank in range(nr):
        P = Problem(problem=1, rank=rank, pgrid=pgrid, **cfg)
        out = (C.c_double * 2)()
        S0, r0 = np.array(P.S0), np.array(P.rho0_gf)
        assert P.lib.lagb_problem_sedov_density_error(P.h, S0.ctypes.data_as(C.c_void_p), r0.ctypes.data_as(C.c_void_p),
                                                      0.2, 1.4, 1.0, 1.0, out) == 0
        tot += out[1]
    assert abs(np.sqrt(tot) - ref) < 1e-12 * ref and ref > 0.1

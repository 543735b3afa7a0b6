"""Independent check of the 1D tables and numbering that the product AND the oracle share
(laghos_b200/csrc/host/fe_tables.hpp, problem.hpp; VERDICT r1 "what's weak" 3).

Nothing here calls fe_tables.hpp's algorithms: Gauss-Legendre comes from numpy's
Golub-Welsch `leggauss`, Gauss-Lobatto points from the roots of P'_n computed by numpy's
companion-matrix root finder, Lagrange values/derivatives from the barycentric formulas, and
Bernstein values from the closed form C(p,l) x^l (1-x)^(p-l).  What the reference takes from MFEM
(SURVEY App. B.1): IntRules.Get(CUBE, 3 ok + ot - 1) (laghos_solver.cpp:145-147), H1 Gauss-Lobatto
Lagrange basis (laghos.cpp:495), positive (Bernstein) L2 basis (laghos.cpp:494).

Closed-form anchors: GLL(4) = {0, (1 -+ 1/sqrt 5)/2, 1}; GLL(5) interior = (1 -+ sqrt(3/7))/2, 1/2;
Gauss(2) = (1 -+ 1/sqrt 3)/2; exactness of the rule for x^k, k <= 2 Q1D - 1.

Also checks the lexicographic gather map / byNODES numbering and the H1 mass operator through
polynomial exactness (independent of the tables' construction): for rho = 1 on the unit cube,
sum(M 1) = 1 and u^t M u = int u^2 for nodal interpolants u of monomials of degree <= ok.
"""
import math

import numpy as np
import pytest
from numpy.polynomial import legendre as L


def gauss01(n):
    x, w = L.leggauss(n)
    return (x + 1) / 2, w / 2


def gll01(n):
    inner = L.Legendre.basis(n - 1).deriv().roots() if n > 2 else np.zeros(0)
    return (np.concatenate([[-1.0], np.sort(inner.real), [1.0]]) + 1) / 2


def lagrange_tables(nodes, xq):
    """Barycentric values and derivatives of the Lagrange basis on `nodes` at `xq` (not nodes)."""
    nodes = nodes.astype(np.longdouble)
    xq = xq.astype(np.longdouble)
    n = len(nodes)
    wb = np.array([1 / np.prod([nodes[j] - nodes[m] for m in range(n) if m != j]) for j in range(n)], dtype=np.longdouble)
    B = np.zeros((n, len(xq)), dtype=np.longdouble)
    G = np.zeros((n, len(xq)), dtype=np.longdouble)
    for q, x in enumerate(xq):
        ell = np.prod(x - nodes)
        for j in range(n):
            B[j, q] = ell * wb[j] / (x - nodes[j])
        for j in range(n):
            # l_j'(x) = l_j(x) * sum_{m != j} 1/(x - x_m)
            G[j, q] = B[j, q] * sum(1 / (x - nodes[m]) for m in range(n) if m != j)
    return B.astype(np.float64), G.astype(np.float64)


@pytest.mark.parametrize("ok", [1, 2, 3, 4, 5])
def test_tables_against_independent_construction(built, ok):
    from laghos_b200.api import Problem
    ot = ok - 1 if ok > 1 else 0
    P = Problem("cube01_hex", 0, 1, ok, ot)
    D, Q, Lk = P.D1D, P.Q1D, P.L1D
    assert D == ok + 1 and Lk == ot + 1
    assert Q == ((3 * ok + ot - 1) | 1) // 2 + 1           # laghos_solver.cpp:145-147 + MFEM's odd-order rule
    qx, qw = gauss01(Q)
    assert np.allclose(P.table(3, Q), qx, rtol=0, atol=2e-15)
    assert np.allclose(P.table(4, Q), qw, rtol=0, atol=2e-15)
    for k in range(2 * Q):                                  # exact for degree <= 2 Q - 1
        assert abs(np.dot(P.table(4, Q), P.table(3, Q) ** k) - 1.0 / (k + 1)) < 5e-15
    nodes = gll01(D)
    B, G = lagrange_tables(nodes, qx)
    Bp = P.table(0, Q * D).reshape(D, Q)
    Gp = P.table(1, Q * D).reshape(D, Q)
    assert np.max(np.abs(Bp - B)) < 5e-14, np.max(np.abs(Bp - B))
    assert np.max(np.abs(Gp - G)) < 5e-13 * max(1.0, np.max(np.abs(G))), np.max(np.abs(Gp - G))
    BL = np.array([[math.comb(ot, l) * x ** l * (1 - x) ** (ot - l) for x in qx] for l in range(Lk)])
    assert np.max(np.abs(P.table(2, Q * Lk).reshape(Lk, Q) - BL)) < 5e-15


def test_closed_form_anchor_points(built):
    from laghos_b200.api import Problem
    s5 = 1 / math.sqrt(5.0)
    assert np.allclose(gll01(4), [0, (1 - s5) / 2, (1 + s5) / 2, 1], atol=1e-15)
    s37 = math.sqrt(3.0 / 7.0)
    assert np.allclose(gll01(5), [0, (1 - s37) / 2, 0.5, (1 + s37) / 2, 1], atol=1e-15)
    # the H1 nodes the product uses are visible through the initial mesh nodes of one element
    for ok, ref in ((3, gll01(4)), (4, gll01(5)), (5, gll01(6))):
        P = Problem("cube01_hex", 0, 1, ok, ok - 1)
        x = P.S0[:P.ndofs_h1]
        n1 = 2 * ok + 1
        row = x[:n1]                                         # first lattice row: x of (gx, 0, 0)
        assert np.allclose(row[:ok + 1], 0.5 * ref, atol=1e-15)
        assert np.allclose(row[ok:], 0.5 + 0.5 * ref, atol=1e-15)


@pytest.mark.parametrize("mesh,rs,ok", [("cube01_hex", 1, 3), ("box01_hex", 0, 2), ("square01_quad", 1, 2)])
def test_gather_map_is_lexicographic(built, mesh, rs, ok):
    from laghos_b200.api import Problem
    P = Problem(mesh, rs, 1, ok, ok - 1)
    dim = P.dim
    coarse = {"cube01_hex": (2, 2, 2), "box01_hex": (4, 2, 2), "square01_quad": (2, 2, 1)}[mesh]
    n = [coarse[d] * 2 ** rs if d < dim else 1 for d in range(3)]
    N1 = [n[d] * ok + 1 if d < dim else 1 for d in range(3)]
    assert P.NE == n[0] * n[1] * n[2] and P.ndofs_h1 == N1[0] * N1[1] * N1[2]
    D = ok + 1
    DZ = D if dim == 3 else 1
    e = np.arange(P.NE)
    ix, iy, iz = e % n[0], (e // n[0]) % n[1], e // (n[0] * n[1])
    kx, ky, kz = np.meshgrid(np.arange(D), np.arange(D), np.arange(DZ), indexing="ij")
    loc = (kx + D * (ky + D * kz)).ravel()
    want = np.zeros((P.NE, P.ND), dtype=np.int64)
    gx = ix[:, None] * ok + kx.ravel()[None, :]
    gy = iy[:, None] * ok + ky.ravel()[None, :]
    gz = iz[:, None] * ok + kz.ravel()[None, :]
    want[:, loc] = gx + N1[0] * (gy + N1[1] * gz)
    assert np.array_equal(P.h1_map.reshape(P.NE, P.ND), want)


@pytest.mark.parametrize("ok", [2, 3, 4, 5])
def test_oracle_mass_polynomial_exactness(built, ok):
    """u^t M u = int_[0,1]^3 rho u^2 with rho = 1 (Sedov) for u = x^a y^b z^c, a,b,c <= ok: pins the
    Q3Q2 / Q4Q3 / Q5Q4 3D tables, weights and map through a closed form, not through shared code."""
    import pyoracle
    from laghos_b200.api import Problem
    P = Problem("cube01_hex", 0, 1, ok, ok - 1)
    O = pyoracle.Oracle("cube01_hex", 0, 1, ok, ok - 1)
    n = P.ndofs_h1
    X = P.S0[:3 * n].reshape(3, n)
    one = np.ones(n)
    assert abs(np.sum(O.vmass_mult(one, -1)) - 1.0) < 1e-13
    for (a, b, c) in [(ok, 0, 0), (1, ok, 0), (ok, ok, ok), (0, 2, ok - 1)]:
        u = X[0] ** a * X[1] ** b * X[2] ** c
        exact = 1.0 / ((2 * a + 1) * (2 * b + 1) * (2 * c + 1))
        got = float(np.dot(u, O.vmass_mult(u, -1)))
        assert abs(got - exact) < 1e-13 * max(1.0, exact) + 1e-15, (a, b, c, got, exact)

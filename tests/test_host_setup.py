"""Host logic (no GPU): tables, numbering, initial conditions, element partition."""
import numpy as np
import pytest


def test_tables_partition_of_unity(built):
    from laghos_b200.api import Problem
    for ok, ot in [(1, 0), (2, 1), (3, 2), (4, 3), (5, 4)]:
        P = Problem("cube01_hex", 0, 1, ok, ot)
        Q, D, L = P.Q1D, P.D1D, P.L1D
        B = P.table(0, Q * D).reshape(D, Q)
        G = P.table(1, Q * D).reshape(D, Q)
        BL = P.table(2, Q * L).reshape(L, Q)
        qw = P.table(4, Q)
        assert np.allclose(B.sum(0), 1.0, atol=1e-14)
        assert np.allclose(G.sum(0), 0.0, atol=1e-12)
        assert np.allclose(BL.sum(0), 1.0, atol=1e-14)
        assert np.all(BL >= 0)
        assert abs(qw.sum() - 1.0) < 1e-15
        assert Q == (3 * ok + ot - 1 | 1) // 2 + 1


def test_sedov_delta_energy(built):
    """sum_q w detJ e(q) = E0/2^dim (DeltaCoefficient scaling, laghos.cpp:603-606)."""
    import pyoracle
    for mesh, rs in [("square01_quad", 2), ("cube01_hex", 1)]:
        O = pyoracle.Oracle(mesh, rs, 1, 2, 1)
        e = O.S0[2 * O.h1_vsize:]
        # unit-density mass operator applied to e and summed = integral of e (rho0 = 1)
        assert abs(O.emass_mult(e).sum() - 1.0 / 2 ** O.dim) < 1e-14


@pytest.mark.parametrize("pgrid", [(2, 1, 1), (2, 2, 1), (2, 2, 2), (3, 1, 2)])
def test_partition_consistency(built, pgrid):
    """Boxes tile the mesh, shared lists pair up, each dof has exactly one owner, and ICs agree."""
    from laghos_b200.api import Problem
    mesh, rs, ok, ot = "cube01_hex", 1, 2, 1
    G = Problem(mesh, rs, 1, ok, ot)
    n = pgrid[0] * pgrid[1] * pgrid[2]
    parts = [Problem(mesh, rs, 1, ok, ot, rank=r, pgrid=pgrid) for r in range(n)]
    assert sum(p.NE for p in parts) == G.NE
    assert sum(int(p.owner_mask.sum()) for p in parts) == G.ndofs_h1
    x_g = G.S0[:G.h1_vsize].reshape(3, -1)
    coords = {tuple(np.round(x_g[:, i], 12)): i for i in range(G.ndofs_h1)}
    for r, p in enumerate(parts):
        xl = p.S0[:p.h1_vsize].reshape(3, -1)
        vl = p.S0[p.h1_vsize:2 * p.h1_vsize].reshape(3, -1)
        vg = G.S0[G.h1_vsize:2 * G.h1_vsize].reshape(3, -1)
        for i in range(0, p.ndofs_h1, 7):
            gi = coords[tuple(np.round(xl[:, i], 12))]
            assert np.array_equal(vl[:, i], vg[:, gi])
        for (nr, ph, dofs) in p.neighbours():
            back = [d for (rr, pp, d) in parts[nr].neighbours() if rr == r and pp == ph]
            assert len(back) == 1 and len(back[0]) == len(dofs)
            xo = parts[nr].S0[:parts[nr].h1_vsize].reshape(3, -1)
            assert np.allclose(xl[:, dofs], xo[:, back[0]], atol=0)
    assert abs(sum(np.abs(p.S0[2 * p.h1_vsize:]).sum() for p in parts) - np.abs(G.S0[2 * G.h1_vsize:]).sum()) < 1e-12


@pytest.mark.parametrize("pgrid", [(2, 2, 2), (2, 2, 1), (3, 1, 2)])
def test_partition_single_phase_sum(built, pgrid):
    """Emulate the single-phase shared-dof exchange (capi.cu halo_pack_all / halo_combine) in numpy for
    every rank of a process grid: each rank adds the values received from ALL its sharers (faces,
    edges, corners) in ascending rank order.  Every copy of a shared dof must end up with the same
    value, equal to the globally assembled sum."""
    from laghos_b200.api import Problem
    mesh, rs, ok, ot = "cube01_hex", 1, 2, 1
    G = Problem(mesh, rs, 1, ok, ot)
    n = pgrid[0] * pgrid[1] * pgrid[2]
    parts = [Problem(mesh, rs, 1, ok, ot, rank=r, pgrid=pgrid) for r in range(n)]
    xg = G.S0[:G.h1_vsize].reshape(3, -1)
    key = {tuple(np.round(xg[:, i], 12)): i for i in range(G.ndofs_h1)}
    l2g = []
    for p in parts:
        xl = p.S0[:p.h1_vsize].reshape(3, -1)
        l2g.append(np.array([key[tuple(np.round(xl[:, i], 12))] for i in range(p.ndofs_h1)]))
    rng = np.random.default_rng(3)
    local = [rng.uniform(-1, 1, p.ndofs_h1) for p in parts]
    assembled = np.zeros(G.ndofs_h1)
    for r in range(n):
        np.add.at(assembled, l2g[r], local[r])
    nbrs = [p.neighbours() for p in parts]
    out = []
    for r in range(n):
        contrib = {}   # dof -> [(rank, value)]
        for (nr, ph, dofs) in nbrs[r]:
            assert ph == 0
            back = [d for (rr, pp, d) in nbrs[nr] if rr == r]
            assert len(back) == 1 and len(back[0]) == len(dofs)
            for mine, theirs in zip(dofs, back[0]):
                contrib.setdefault(int(mine), []).append((nr, local[nr][theirs]))
        v = local[r].copy()
        for dof, lst in contrib.items():
            lst.append((r, local[r][dof]))
            acc = 0.0
            for _, val in sorted(lst):
                acc += val
            v[dof] = acc
        out.append(v)
    for r in range(n):
        assert np.allclose(out[r], assembled[l2g[r]], rtol=0, atol=1e-14)
    # bit-identical copies of every shared dof across ranks (rank-ordered summation)
    owner_val = {}
    for r in range(n):
        for i, g in enumerate(l2g[r]):
            if g in owner_val:
                assert owner_val[g] == out[r][i]
            else:
                owner_val[g] = out[r][i]
    assert sum(int(p.owner_mask.sum()) for p in parts) == G.ndofs_h1


def test_oracle_energies(built):
    """InternalEnergy / KineticEnergy (reference laghos_solver.cpp:639-697): the Sedov blast energy is
    E0/2^dim, and the integrals equal 1^t M_L2 e and 1/2 v^t M_H1 v (the form the CUDA path evaluates)."""
    import pyoracle
    O = pyoracle.Oracle("cube01_hex", 1, 1, 2, 1)
    e = O.S0[2 * O.h1_vsize:]
    assert abs(O.internal_energy(e) - 0.125) < 1e-15
    assert abs(O.internal_energy(e) - O.emass_mult(e).sum()) < 1e-15
    v = np.random.default_rng(0).uniform(-1, 1, O.h1_vsize)
    n = O.ndofs_h1
    ke = 0.5 * sum(float(v[c * n:(c + 1) * n] @ O.vmass_mult(np.ascontiguousarray(v[c * n:(c + 1) * n]))) for c in range(3))
    assert abs(O.kinetic_energy(v) - ke) < 1e-14 * ke
    r = pyoracle.run(mesh="cube01_hex", rs=1, problem=0, ok=2, ot=1, max_tsteps=10, t_final=1e9, cg_tol=1e-12)
    assert abs(r["energy_init"] - r["energy_final"]) < 1e-9 * r["energy_init"]   # smooth Taylor-Green: RK4 conserves to ~1e-11


# ---- brick schedule of the mass apply (host/batch_plan.hpp) ----
@pytest.mark.parametrize("mesh,rs,ok,grid_hint,nb", [
    ("cube01_hex", 2, 3, True, 8), ("cube01_hex", 2, 3, True, 16), ("cube01_hex", 1, 2, True, 16),
    ("box01_hex", 1, 3, True, 8), ("cube01_hex", 2, 3, False, 8), ("cube01_hex", 0, 4, True, 8),
    ("cube01_hex", 1, 1, False, 32)])
def test_batch_plan_invariants(built, mesh, rs, ok, grid_hint, nb):
    """Every element scheduled once, colours conflict-free, one first writer per dof (checked in C++),
    plus the counts a Cartesian brick must have."""
    import ctypes as C
    from laghos_b200.api import Problem
    P = Problem(mesh, rs, 1, ok, ok - 1)
    mp = np.ascontiguousarray(P.h1_map, dtype=np.int32)
    ne = [(int(P.info.n1[k]) - 1) // (P.D1D - 1) for k in range(3)]
    grid = (C.c_int32 * 3)(*(ne if grid_hint else [0, 0, 0]))
    stats = (C.c_int64 * 8)()
    rc = P.lib.lagb_host_batch_plan_check(mp.ctypes.data_as(C.POINTER(C.c_int32)), P.NE, P.ND, P.ndofs_h1, grid, nb, stats)
    assert rc == 0, P.lib.lagb_last_error().decode()
    nbatch, ncolors, ntab, umax, UP, nfirst, shape, tot = [int(v) for v in stats]
    assert nfirst == P.ndofs_h1 and UP >= umax and UP % 32 == 0
    if grid_hint:
        b = [shape % 100, (shape // 100) % 100, shape // 10000]
        assert b[0] * b[1] * b[2] <= nb
        full = all(ne[k] % b[k] == 0 for k in range(3))
        if full:
            assert umax == np.prod([b[k] * ok + 1 for k in range(3)])
            assert nbatch == np.prod([ne[k] // b[k] for k in range(3)])
            assert ntab == 1            # all bricks share one index table
        assert ncolors <= 8
    else:
        assert nbatch == (P.NE + nb - 1) // nb


def test_batch_plan_apply_emulation(built):
    """numpy emulation of the coloured schedule: first writers store, later colours add; the result
    equals the plain E^t E sum and no zero fill is needed."""
    from laghos_b200.api import Problem
    P = Problem("cube01_hex", 1, 1, 3, 2)
    mp = P.h1_map.reshape(P.NE, P.ND)
    rng = np.random.default_rng(5)
    ev = rng.uniform(-1, 1, (P.NE, P.ND))
    ref = np.zeros(P.ndofs_h1)
    np.add.at(ref, mp.ravel(), ev.ravel())
    # bricks 2x2x2 with parity colours, as BatchPlan builds them
    n = [(int(P.info.n1[k]) - 1) // (P.D1D - 1) for k in range(3)]
    y = np.full(P.ndofs_h1, np.nan)      # garbage: the schedule never reads before the first writer stored
    bricks = {}
    for e in range(P.NE):
        ex, ey, ez = e % n[0], (e // n[0]) % n[1], e // (n[0] * n[1])
        kb = (ex // 2, ey // 2, ez // 2)
        bricks.setdefault(kb, []).append(e)
    mincol = {}
    for kb, els in bricks.items():
        col = (kb[0] & 1) | ((kb[1] & 1) << 1) | ((kb[2] & 1) << 2)
        for d in np.unique(mp[els]):
            mincol[d] = min(mincol.get(d, 99), col)
    for col in range(8):
        for kb, els in bricks.items():
            if ((kb[0] & 1) | ((kb[1] & 1) << 1) | ((kb[2] & 1) << 2)) != col:
                continue
            u, inv = np.unique(mp[els].ravel(), return_inverse=True)
            s = np.zeros(len(u))
            np.add.at(s, inv, ev[els].ravel())
            for j, d in enumerate(u):
                if mincol[d] == col:
                    y[d] = s[j]
                else:
                    y[d] += s[j]
    assert np.all(np.isfinite(y))
    assert np.max(np.abs(y - ref)) < 1e-14


# ---- 1D tables against closed forms that do not share code with host/fe_tables.hpp ----
def test_tables_closed_form(built):
    """Gauss-Legendre points/weights (numpy), Gauss-Lobatto nodes (roots of (1-x^2) P'_{p}), nodal
    Lagrange values/derivatives (numpy polynomial arithmetic) and Bernstein values for orders 1-5."""
    from numpy.polynomial import legendre as L, polynomial as Pn
    from math import comb
    from laghos_b200.api import Problem
    for ok, ot in [(1, 0), (2, 1), (3, 2), (4, 3), (5, 4)]:
        P = Problem("cube01_hex", 0, 1, ok, ot)
        Q, D, Lb = P.Q1D, P.D1D, P.L1D
        qx, qw = P.table(3, Q), P.table(4, Q)
        gx, gw = L.leggauss(Q)
        assert np.max(np.abs(qx - 0.5 * (gx + 1))) < 1e-15 and np.max(np.abs(qw - 0.5 * gw)) < 1e-15
        # GLL nodes on [0,1]
        if ok == 1:
            nodes = np.array([0.0, 1.0])
        else:
            c = np.zeros(ok + 1); c[ok] = 1.0
            inner = np.sort(L.legroots(L.legder(c)))
            nodes = 0.5 * (np.concatenate([[-1.0], inner, [1.0]]) + 1)
        B = P.table(0, Q * D).reshape(D, Q)
        G = P.table(1, Q * D).reshape(D, Q)
        for d in range(D):
            others = np.delete(nodes, d)
            den = np.prod(nodes[d] - others)
            val = np.array([np.prod(x - others) for x in qx]) / den            # product form (well conditioned)
            der = np.array([sum(np.prod(x - np.delete(others, k)) for k in range(len(others))) for x in qx]) / den
            assert np.max(np.abs(B[d] - val)) < 5e-15
            assert np.max(np.abs(G[d] - der)) < 2e-13
        BL = P.table(2, Q * Lb).reshape(Lb, Q)
        for l in range(Lb):
            assert np.max(np.abs(BL[l] - comb(ot, l) * qx ** l * (1 - qx) ** (ot - l))) < 1e-15

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


def test_quad_transpose_emulation():
    """Index logic of the experimental shuffle hand-off (device/mass3d_shfl.cuh, quad_transpose): a two-step
    xor butterfly over 4 lanes with compile-time register indices transposes the 4 x 4 block matrix
    (lane j, slot m) <- (lane m, column 4k + j) and is an involution.  Emulated lane by lane."""
    NK = 9
    V = np.array([[100.0 * j + col for col in range(4 * NK)] for j in range(4)])

    def transpose(V):
        V = V.copy()
        for p in (0, 2):
            for k in range(NK):
                send = [V[j, 4 * k + p] if (j & 1) else V[j, 4 * k + p + 1] for j in range(4)]
                for j in range(4):
                    V[j, 4 * k + p + (0 if j & 1 else 1)] = send[j ^ 1]
        for s in (0, 1):
            for k in range(NK):
                send = [V[j, 4 * k + s] if (j & 2) else V[j, 4 * k + 2 + s] for j in range(4)]
                for j in range(4):
                    V[j, 4 * k + (s if j & 2 else 2 + s)] = send[j ^ 2]
        return V

    T = transpose(V)
    assert all(T[j, 4 * k + m] == 100.0 * m + 4 * k + j for j in range(4) for k in range(NK) for m in range(4))
    assert np.array_equal(transpose(T), V)

"""Output files (reference `-print`, laghos.cpp:873-900, and `-visit`, :866-871): MFEM mesh v1.0 with
a `nodes` grid function + GridFunction::Save files written by the host side (host/mesh_writer.hpp) through the C ABI.
Checked here with an independent parser: topology (right-handed elements, outward boundary faces with the reference's
attribute convention), exact field values at full precision, the round trip through the mesh reader, the VisIt root
file, a partitioned block, and the error path."""
import json
import os

import numpy as np
import pytest

from laghos_b200.api import LagbError, Problem


def parse_mesh(path):
    lines = [ln.split("#")[0].strip() for ln in open(path)]
    assert open(path).readline().startswith("MFEM mesh v1.0")
    tok = " ".join(lines[1:]).split()
    p = tok.index("dimension")
    dim = int(tok[p + 1])
    p = tok.index("elements")
    ne = int(tok[p + 1])
    p += 2
    nve = 2 ** dim
    elems = []
    for _ in range(ne):
        attr, geom = int(tok[p]), int(tok[p + 1])
        assert attr == 1 and geom == (3 if dim == 2 else 5)
        elems.append([int(t) for t in tok[p + 2:p + 2 + nve]])
        p += 2 + nve
    assert tok[p] == "boundary"
    nb = int(tok[p + 1])
    p += 2
    nvb = 2 ** (dim - 1)
    bnd = []
    for _ in range(nb):
        attr, geom = int(tok[p]), int(tok[p + 1])
        assert geom == (1 if dim == 2 else 3)
        bnd.append((attr, [int(t) for t in tok[p + 2:p + 2 + nvb]]))
        p += 2 + nvb
    assert tok[p] == "vertices"
    nv = int(tok[p + 1])
    assert tok[p + 2:p + 4] == ["nodes", "FiniteElementSpace"]
    fec, vdim, ordering, vals = parse_gf_tokens(tok[p + 3:])
    return dict(dim=dim, elems=np.array(elems), bnd=bnd, nv=nv, fec=fec, vdim=vdim, ordering=ordering, nodes=vals)


def parse_gf_tokens(tok):
    assert tok[0] == "FiniteElementSpace" and tok[1] == "FiniteElementCollection:"
    fec = tok[2]
    assert tok[3] == "VDim:" and tok[5] == "Ordering:"
    return fec, int(tok[4]), int(tok[6]), np.array([float(t) for t in tok[7:]])


def parse_gf(path):
    return parse_gf_tokens(open(path).read().split())


CASES = [("cube01_hex", 1, 1, 3, 2), ("square01_quad", 1, 0, 2, 1), ("box01_hex", 0, 3, 2, 1), ("rectangle01_quad", 1, 3, 4, 3)]


@pytest.mark.parametrize("mesh,rs,problem,ok,ot", CASES)
def test_mesh_file_topology_and_nodes(built, tmp_path, mesh, rs, problem, ok, ot):
    P = Problem(mesh=mesh, rs=rs, problem=problem, ok=ok, ot=ot)
    dim, NE, ND, D = P.dim, P.NE, P.ND, P.D1D
    x = P.S0[:dim * P.ndofs_h1].copy()
    # a smooth deformation: the file must carry the CURRENT positions, not the initial ones
    x += 1e-3 * np.sin(np.arange(x.size))
    path = tmp_path / "m.mesh"
    P.write_mesh(path, x, precision=17)
    M = parse_mesh(path)
    assert M["dim"] == dim and len(M["elems"]) == NE
    assert M["fec"] == f"L2_T1_{dim}D_P{ok}" and M["vdim"] == dim and M["ordering"] == 0
    nodes = M["nodes"].reshape(dim, NE * ND)
    expect = x.reshape(dim, -1)[:, P.h1_map]
    assert np.array_equal(nodes, expect)                      # 17 digits: exact round trip of every double
    # corner nodes of every element, MFEM corner order
    cq = [(0, 0), (1, 0), (1, 1), (0, 1)]
    corner_loc = [(D - 1) * cq[c % 4][0] + D * ((D - 1) * cq[c % 4][1] + D * (D - 1) * (c // 4)) for c in range(2 ** dim)]
    X0 = P.S0[:dim * P.ndofs_h1].reshape(dim, -1)[:, P.h1_map].reshape(dim, NE, ND)[:, :, corner_loc]   # undeformed
    V = np.full((M["nv"], dim), np.nan)
    for e in range(NE):
        for c, v in enumerate(M["elems"][e]):
            xc = X0[:, e, c]
            assert np.isnan(V[v, 0]) or np.allclose(V[v], xc, atol=1e-14)    # shared vertices: one position
            V[v] = xc
    assert not np.isnan(V).any()                               # every vertex is used
    assert len(np.unique(np.round(V, 12), axis=0)) == M["nv"]  # and distinct
    # right-handed elements
    for e in range(NE):
        v = V[M["elems"][e]]
        if dim == 2:
            a, b = v[1] - v[0], v[3] - v[0]
            assert a[0] * b[1] - a[1] * b[0] > 0
        else:
            assert np.dot(np.cross(v[1] - v[0], v[3] - v[0]), v[4] - v[0]) > 0
    # boundary: attribute k on faces of constant x_{k-1} at the domain ends, outward orientation, complete cover
    lo, hi = V.min(axis=0), V.max(axis=0)
    n_axis = [len(np.unique(np.round(V[:, a], 12))) - 1 for a in range(dim)]
    expect_nb = sum(2 * int(np.prod([n_axis[b] for b in range(dim) if b != a])) for a in range(dim))
    assert len(M["bnd"]) == expect_nb
    seen = set()
    for attr, vs in M["bnd"]:
        a = attr - 1
        v = V[vs]
        assert np.ptp(v[:, a]) < 1e-14
        side = 1 if abs(v[0, a] - hi[a]) < 1e-14 else 0
        assert side == 1 or abs(v[0, a] - lo[a]) < 1e-14
        if dim == 2:
            t = v[1] - v[0]
            normal = np.array([t[1], -t[0]])                   # domain on the left of the edge -> outward = right
        else:
            normal = np.cross(v[1] - v[0], v[3] - v[0])
        assert normal[a] * (1 if side else -1) > 0
        key = tuple(sorted(vs))
        assert key not in seen
        seen.add(key)


@pytest.mark.parametrize("mesh,rs,problem,ok,ot", CASES[:3])
def test_written_initial_mesh_reads_back(built, tmp_path, mesh, rs, problem, ok, ot):
    P = Problem(mesh=mesh, rs=rs, problem=problem, ok=ok, ot=ot)
    path = tmp_path / "m0.mesh"
    P.write_mesh(path, None, precision=17)
    Q = Problem(mesh_file=path, rs=0, problem=problem, ok=ok, ot=ot, dim=P.dim)
    assert (Q.NE, Q.ndofs_h1, Q.ndofs_l2) == (P.NE, P.ndofs_h1, P.ndofs_l2)
    assert np.array_equal(Q.h1_map, P.h1_map)
    assert np.allclose(Q.S0, P.S0, rtol=0, atol=1e-14)
    for a in range(P.dim):
        assert np.allclose(Q.mesh_breaks(a), P.mesh_breaks(a), rtol=0, atol=1e-15)


def test_print_files(built, tmp_path):
    P = Problem(mesh="cube01_hex", rs=1, problem=1, ok=2, ot=1)
    rng = np.random.default_rng(3)
    S = P.S0 + 1e-3 * rng.standard_normal(P.s_size)
    rho = rng.random(P.ndofs_l2)
    base = str(tmp_path / "results" / "Laghos")
    os.makedirs(os.path.dirname(base))
    P.write_print(base, 5, S, rho, precision=17)
    nv = P.dim * P.ndofs_h1
    M = parse_mesh(base + "_5_mesh")
    assert np.array_equal(M["nodes"].reshape(3, -1), S[:nv].reshape(3, -1)[:, P.h1_map])
    fec, vdim, ordering, v = parse_gf(base + "_5_v")
    assert (fec, vdim, ordering) == ("L2_T1_3D_P2", 3, 0)
    assert np.array_equal(v.reshape(3, -1), S[nv:2 * nv].reshape(3, -1)[:, P.h1_map])
    fec, vdim, ordering, e = parse_gf(base + "_5_e")
    assert (fec, vdim, ordering) == ("L2_T2_3D_P1", 1, 0) and np.array_equal(e, S[2 * nv:])
    fec, vdim, ordering, r = parse_gf(base + "_5_rho")
    assert (fec, vdim, ordering) == ("L2_T2_3D_P1", 1, 0) and np.array_equal(r, rho)
    # the reference's default stream precision (laghos.cpp:882 `precision(8)`): 8 significant digits
    P.write_print(base, 6, S, rho)
    _, _, _, e8 = parse_gf(base + "_6_e")
    assert np.allclose(e8, S[2 * nv:], rtol=1e-7, atol=0) and not np.array_equal(e8, S[2 * nv:])


def test_visit_collection(built, tmp_path):
    P = Problem(mesh="square01_quad", rs=1, problem=0, ok=2, ot=1)
    S = np.array(P.S0)
    os.makedirs(tmp_path / "results")
    coll = str(tmp_path / "results" / "Laghos")             # the reference's default basename "results/Laghos"
    P.write_visit(coll, 7, 0.25, 0.01, S, rho=np.ones(P.ndofs_l2))
    root = json.load(open(coll + "_000007.mfem_root"))["dsets"]["main"]
    assert root["cycle"] == 7 and root["domains"] == 1 and root["time"] == 0.25 and root["time_step"] == 0.01
    assert root["mesh"]["tags"]["spatial_dim"] == "2" and root["mesh"]["format"] == "0"
    assert set(root["fields"]) == {"Density", "Velocity", "Specific Internal Energy"}   # laghos.cpp:695-697
    # paths are relative to the root file's directory (MFEM DataCollection: prefix path + name)
    assert root["mesh"]["path"] == "Laghos_000007/mesh.%06d"
    here = os.path.dirname(coll)
    assert os.path.isfile(os.path.join(here, root["mesh"]["path"] % 0))
    for name, f in root["fields"].items():
        fec, vdim, _, vals = parse_gf(os.path.join(here, f["path"] % 0))
        assert int(f["tags"]["comps"]) == vdim and f["tags"]["assoc"] == "nodes"
        assert f["tags"]["lod"] == ("2" if name == "Velocity" else "1")
        assert fec.startswith("L2_T1_2D_P2" if name == "Velocity" else "L2_T2_2D_P1")
    # a second cycle into a new directory, without a density field; rank 1 of 2 writes no root file
    P.write_visit(coll, 8, 0.3, 0.01, S, rank=1, nranks=2)
    assert os.path.isfile(coll + "_000008/mesh.000001") and not os.path.exists(coll + "_000008.mfem_root")
    assert not os.path.exists(coll + "_000008/Density.000001")


def test_partitioned_block(built, tmp_path):
    # rank 1 of a 2 x 1 x 1 grid: the cut face x = 0.5 is not a boundary of the global domain
    P = Problem(mesh="cube01_hex", rs=1, problem=1, ok=2, ot=1, rank=1, pgrid=(2, 1, 1))
    path = tmp_path / "part.mesh"
    P.write_mesh(path, None, precision=17)
    M = parse_mesh(path)
    assert len(M["elems"]) == P.NE == 32
    attrs = [a for a, _ in M["bnd"]]
    assert attrs.count(1) == 16 and attrs.count(2) == 2 * 8 and attrs.count(3) == 2 * 8
    x = M["nodes"].reshape(3, -1)[0]
    assert x.min() == 0.5 and x.max() == 1.0


def test_write_errors(built, tmp_path):
    P = Problem(mesh="square01_quad", rs=0, problem=0, ok=2, ot=1)
    with pytest.raises(LagbError, match="cannot open"):
        P.write_mesh(tmp_path / "no_such_dir" / "m.mesh")
    with pytest.raises(LagbError, match="kind"):
        P.write_field(tmp_path / "f", np.zeros(P.ndofs_l2), kind=2)
    with pytest.raises(ValueError):
        P.write_field(tmp_path / "f", np.zeros(3), kind=1)

"""Mesh-file input (reference `-m data/<mesh>.mesh`, MFEM mesh v1.0): the rectilinear subset the host
setup supports.  (1) Files written by the test in both vertex encodings reproduce the generated
problem; (2) when the reference tree is present (this container, not the GPU box) its own data/ files
must give exactly the breakpoints that `named_coarse_mesh` hard-codes; (3) unsupported input is rejected."""
import itertools
import os

import numpy as np
import pytest

REF_DATA = "/root/reference/data"


def write_mfem_mesh(path, breaks, nodes_block=False, bad_attr=False):
    dim = len(breaks)
    n = [len(b) - 1 for b in breaks]
    nv1 = [len(b) for b in breaks]

    def vid(idx):
        return idx[0] + nv1[0] * (idx[1] + (nv1[1] * idx[2] if dim == 3 else 0))

    lines = ["MFEM mesh v1.0", "", "# written by tests/test_mesh_reader.py", "", "dimension", str(dim), ""]
    elems = []
    for cell in itertools.product(*[range(k) for k in reversed(n)]):
        c = cell[::-1]
        if dim == 2:
            vs = [vid((c[0], c[1], 0)), vid((c[0] + 1, c[1], 0)), vid((c[0] + 1, c[1] + 1, 0)), vid((c[0], c[1] + 1, 0))]
            elems.append("1 3 " + " ".join(map(str, vs)))
        else:
            vs = [vid((c[0] + a, c[1] + b, c[2] + d)) for d in (0, 1) for (a, b) in ((0, 0), (1, 0), (1, 1), (0, 1))]
            elems.append("1 5 " + " ".join(map(str, vs)))
    lines += ["elements", str(len(elems))] + elems + [""]
    bnd = []
    for axis in range(dim):
        for side in (0, n[axis]):
            others = [a for a in range(dim) if a != axis]
            for cell in itertools.product(*[range(n[a]) for a in others]):
                corners = []
                for off in itertools.product(*[(0, 1)] * len(others)):
                    idx = [0, 0, 0]
                    idx[axis] = side
                    for a, c0, o in zip(others, cell, off):
                        idx[a] = c0 + o
                    corners.append(vid(tuple(idx)))
                if dim == 3:
                    corners = [corners[0], corners[1], corners[3], corners[2]]
                attr = axis + 1 if not bad_attr else 1
                bnd.append(f"{attr} {1 if dim == 2 else 3} " + " ".join(map(str, corners)))
    lines += ["boundary", str(len(bnd))] + bnd + [""]
    nv = int(np.prod(nv1))
    coords = np.zeros((nv, dim))
    for idx in itertools.product(*[range(k) for k in nv1]):
        full = tuple(idx) + (0,) * (3 - dim)
        coords[vid(full)] = [breaks[a][idx[a]] for a in range(dim)]
    if nodes_block:
        lines += ["vertices", str(nv), "", "nodes", "FiniteElementSpace", "FiniteElementCollection: Linear",
                  f"VDim: {dim}", "Ordering: 0", ""]
        for a in range(dim):
            lines += [repr(float(v)) for v in coords[:, a]]
    else:
        lines += ["vertices", str(nv), str(dim)] + [" ".join(repr(float(v)) for v in row) for row in coords]
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")


@pytest.mark.parametrize("nodes_block", [False, True])
@pytest.mark.parametrize("breaks,name,problem", [
    ([[0, .5, 1], [0, .5, 1], [0, .5, 1]], "cube01_hex", 1),
    ([[0, 1, 3, 5, 7], [0, 1.5, 3], [0, 1.5, 3]], "box01_hex", 3),
    ([[0, .5, 1], [0, .5, 1]], "square01_quad", 0),
])
def test_written_file_reproduces_named_problem(built, tmp_path, breaks, name, problem, nodes_block):
    from laghos_b200.api import Problem
    path = tmp_path / f"{name}.mesh"
    write_mfem_mesh(path, breaks, nodes_block=nodes_block)
    A = Problem(name, 1, problem, 2, 1)
    B = Problem(rs=1, problem=problem, ok=2, ot=1, mesh_file=path, dim=len(breaks))
    assert (A.dim, A.NE, A.ndofs_h1, A.ndofs_l2) == (B.dim, B.NE, B.ndofs_h1, B.ndofs_l2)
    assert np.array_equal(A.h1_map, B.h1_map)
    assert np.array_equal(A.S0, B.S0) and np.array_equal(A.gamma, B.gamma) and np.array_equal(A.rho0_gf, B.rho0_gf)
    for c in range(A.dim):
        assert np.array_equal(A.ess(c), B.ess(c))


def test_rejects_unsupported(built, tmp_path):
    from laghos_b200.api import Problem, LagbError
    p = tmp_path / "bad_attr.mesh"
    write_mfem_mesh(p, [[0, .5, 1], [0, .5, 1]], bad_attr=True)
    with pytest.raises(LagbError, match="boundary attributes"):
        Problem(rs=0, problem=0, mesh_file=p, dim=2)
    q = tmp_path / "tri.mesh"
    q.write_text("MFEM mesh v1.0\n\ndimension\n2\n\nelements\n1\n1 2 0 1 2\n\nboundary\n0\n\nvertices\n3\n2\n0 0\n1 0\n0 1\n")
    with pytest.raises(LagbError, match="tensor-product"):
        Problem(rs=0, problem=0, mesh_file=q, dim=2)
    with pytest.raises(LagbError, match="cannot open"):
        Problem(rs=0, problem=0, mesh_file=tmp_path / "missing.mesh", dim=2)


@pytest.mark.skipif(not os.path.isdir(REF_DATA), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("name,dim", [("cube01_hex", 3), ("box01_hex", 3), ("square01_quad", 2),
                                      ("rectangle01_quad", 2), ("square_gresho", 2), ("rt2D", 2)])
def test_reference_data_files_match_named_meshes(built, name, dim):
    """The hard-coded breakpoints of named_coarse_mesh (host/problem.hpp) against the reference's files."""
    from laghos_b200.api import Problem
    A = Problem(name, 0, 1 if dim == 3 else 0, 2, 1)
    B = Problem(rs=0, problem=1 if dim == 3 else 0, ok=2, ot=1, mesh_file=os.path.join(REF_DATA, name + ".mesh"), dim=dim)
    for a in range(dim):
        assert np.array_equal(A.mesh_breaks(a), B.mesh_breaks(a)), (name, a, A.mesh_breaks(a), B.mesh_breaks(a))
    assert np.array_equal(A.S0, B.S0) and np.array_equal(A.h1_map, B.h1_map)


def test_builtin_default_mesh(built):
    """reference `-m default` (laghos.cpp:131-137, 427-447): MakeCartesian(nx, ny, nz; Sx, Sy, Sz), the mesh of --checks"""
    from laghos_b200.api import LagbError, Problem, mesh_dim
    a, b = Problem(mesh="default", rs=1, problem=1), Problem(mesh="cube01_hex", rs=1, problem=1)
    assert np.array_equal(a.S0, b.S0) and np.array_equal(a.h1_map, b.h1_map)
    a, b = Problem(mesh="default_2d", rs=2, problem=0), Problem(mesh="square01_quad", rs=2, problem=0)
    assert a.dim == 2 and np.array_equal(a.S0, b.S0)
    c = Problem(mesh="default_4x3x2_S2x1.5x0.5", rs=0, problem=1)
    assert c.dim == 3 and c.NE == 24
    assert np.allclose(c.mesh_breaks(0), np.linspace(0, 2, 5)) and np.allclose(c.mesh_breaks(1), np.linspace(0, 1.5, 4))
    assert np.allclose(c.mesh_breaks(2), [0, 0.25, 0.5])
    d = Problem(mesh="default_3x5", rs=0, problem=0)
    assert d.dim == 2 and d.NE == 15 and mesh_dim("default_3x5_S2x1") == 2
    assert np.allclose(Problem(mesh="default_3x5_S2x1", rs=0, problem=0).mesh_breaks(0), [0, 2 / 3, 4 / 3, 2])
    for bad in ("default_0x2", "default_2x2_S0x1", "default_2x2x2_Sax1x1", "defaultx"):
        with pytest.raises(LagbError):
            Problem(mesh=bad, rs=0, problem=1, dim=3)

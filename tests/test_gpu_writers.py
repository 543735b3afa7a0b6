"""`-print` / `-visit` output of the driver loop on the GPU (reference laghos.cpp:845-900): at every vis_steps-th
accepted step and at the last one the density is projected on the current mesh (device) and mesh + rho + v + e go
to files through host/mesh_writer.hpp.  The files of the last step must hold the returned end state."""
import json
import os

import numpy as np
import pytest

from test_mesh_writer import parse_gf, parse_mesh

pytestmark = pytest.mark.gpu


def test_print_and_visit_files_of_a_run(built, tmp_path):
    from laghos_b200.api import Problem, run
    cfg = dict(mesh="cube01_hex", rs=1, problem=1, ok=2, ot=1)
    base = str(tmp_path / "Laghos")
    r = run(**cfg, t_final=10.0, max_tsteps=3, vis_steps=1, gfprint=True, visit=True, basename=base, want_state=True,
            hist_cap=64)
    P = Problem(**cfg)
    S, nv, last = r["S"], P.dim * P.ndofs_h1, r["ti_last"]
    written = sorted(int(f.split("_")[-2]) for f in os.listdir(tmp_path) if f.endswith("_mesh"))
    assert written == [ti for ti, _ in r["hist"]] and written[-1] == last      # every accepted step (vis_steps 1)
    M = parse_mesh(f"{base}_{last}_mesh")
    assert M["fec"] == "L2_T1_3D_P2" and len(M["elems"]) == P.NE
    assert np.allclose(M["nodes"].reshape(3, -1), S[:nv].reshape(3, -1)[:, P.h1_map], rtol=1e-7, atol=1e-12)
    _, vdim, _, v = parse_gf(f"{base}_{last}_v")
    assert vdim == 3 and np.allclose(v.reshape(3, -1), S[nv:2 * nv].reshape(3, -1)[:, P.h1_map], rtol=1e-7, atol=1e-12)
    fec, _, _, e = parse_gf(f"{base}_{last}_e")
    assert fec == "L2_T2_3D_P1" and np.allclose(e, S[2 * nv:], rtol=1e-7, atol=1e-12)
    _, _, _, rho = parse_gf(f"{base}_{last}_rho")
    assert rho.size == P.ndofs_l2 and np.all(np.isfinite(rho)) and rho.min() > 0.0
    # mass is conserved: sum over the L2 dofs of the projected density stays near rho0 = 1 on average
    assert abs(rho.mean() - 1.0) < 0.2
    root = json.load(open(f"{base}_{last:06d}.mfem_root"))["dsets"]["main"]
    assert root["cycle"] == last and root["domains"] == 1 and abs(root["time"] - r["t"]) < 1e-12
    here = os.path.dirname(base)                      # paths in the root file are relative to its directory
    for f in list(root["fields"].values()) + [root["mesh"]]:
        assert os.path.isfile(os.path.join(here, f["path"] % 0))
    _, _, _, e2 = parse_gf(os.path.join(here, root["fields"]["Specific Internal Energy"]["path"] % 0))
    assert np.array_equal(e2, e)


def test_velocity_error_of_a_run(built):
    """laghos.cpp:970-982: the driver's L_inf / L_1 / L_2 velocity errors (problems 0 and 4) are those of the end state"""
    from laghos_b200.api import Problem, run
    cfg = dict(mesh="square01_quad", rs=2, problem=0, ok=2, ot=1)
    r = run(**cfg, t_final=10.0, max_tsteps=5, v_error=True, want_state=True)
    assert np.allclose(r["v_err"], Problem(**cfg).velocity_error(r["S"]), rtol=1e-14)
    assert 0 < r["v_err"][1] < r["v_err"][2] < r["v_err"][0] < 0.1     # smooth flow, a few steps: small errors
    import pyoracle
    ro = pyoracle.run(**cfg, t_final=10.0, max_tsteps=5, want_state=True)             # CPU oracle: same end state
    assert np.allclose(r["v_err"], Problem(**cfg).velocity_error(ro["S"]), rtol=1e-7)
    assert run(**dict(cfg, problem=1), t_final=10.0, max_tsteps=2, v_error=True)["v_err"] == [0.0, 0.0, 0.0]


def test_sedov_density_error_of_a_run(built):
    """-err (laghos.cpp:1009-1085): "Density L2 error" of the driver = the host formula on the end state, with the
    density projected by the CPU oracle's ComputeDensity; and the guard against shock reflections"""
    import pyoracle
    from laghos_b200.api import Problem, run
    cfg = dict(mesh="cube01_hex", rs=1, problem=1, ok=2, ot=1)
    r = run(**cfg, t_final=0.05, check_exact_sedov=True, want_state=True)
    assert abs(r["t"] - 0.05) < 1e-14 and r["density_l2_err"] > 0
    P = Problem(**cfg)
    rho = pyoracle.Oracle(**cfg).compute_density(r["S"][:P.dim * P.ndofs_h1])
    assert abs(r["density_l2_err"] - P.sedov_density_error(r["S"], rho, r["t"])) < 1e-8 * r["density_l2_err"]
    # a coarse mesh smears the shock: the error is O(1) times sqrt(shock volume), far below the ambient norm 1
    assert 0.01 < r["density_l2_err"] < 1.0
    assert run(**cfg, t_final=0.01)["density_l2_err"] == 0.0


def test_checks_mode_of_a_run(built):
    """--checks (laghos.cpp:904-926): the driver loop compares |e| with the reference's table itself.  The B200 path
    sums in a different order than the reference (atomics, batched PCG): gate 1e-11 as in test_gpu_end_to_end.py."""
    from laghos_b200.api import LagbError, run
    kw = dict(rs=0, ok=2, ot=1, t_final=0.6, cfl=0.5, cg_tol=1e-14, check=True, check_eps=1e-11)
    assert run(mesh="square01_quad", problem=1, **kw)["checks"] == 2
    assert run(mesh="cube01_hex", problem=1, **kw)["checks"] == 2
    with pytest.raises(LagbError, match="check: rs, rp"):
        run(mesh="square01_quad", problem=1, **dict(kw, rs=1))
    with pytest.raises(LagbError, match="check failed: P1, #5"):      # a loose CG moves |e| by far more than 1e-11
        run(mesh="square01_quad", problem=1, **dict(kw, cg_tol=1e-4))

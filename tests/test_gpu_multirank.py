"""Multi-GPU parity (needs >= 2 GPUs; skipped otherwise): an element-partitioned run over NCCL
(shared-dof sums, owner-masked CG dots, min dt) reproduces the single-GPU run of the same global
mesh: identical step sequence, |e| within 1e-9 (the reference's criterion is that
the same --checks table passes for every rank count, laghos.cpp:1441-1463)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


# (2,2,1): edge dofs shared by 4 ranks; (2,2,2): the 26-neighbour single-phase exchange of the 8-GPU
# weak-scaling run (faces, edges shared by 4 ranks, the centre corner shared by all 8)
@pytest.mark.parametrize("pgrid,problem", [((2, 1, 1), 1), ((1, 2, 1), 0), ((2, 2, 1), 1), ((2, 2, 2), 1), ((2, 2, 2), 0)],
                         ids=["2x1x1-sedov", "1x2x1-tg", "2x2x1-sedov", "2x2x2-sedov", "2x2x2-tg"])
def test_n_gpu_matches_one(built, pgrid, problem):
    nr = pgrid[0] * pgrid[1] * pgrid[2]
    if _ngpu() < nr:
        pytest.skip(f"needs {nr} GPUs")
    from laghos_b200.api import run
    kw = dict(mesh="cube01_hex", rs=2, problem=problem, ok=3, ot=2, max_tsteps=6, t_final=1e9, cg_tol=1e-12)
    ref = run(**kw, hist_cap=1024)
    port = 29600 + os.getpid() % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nr),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "mgpu_check.py"),
           "--pgrid", ",".join(map(str, pgrid)), "--rs", "2", "--problem", str(problem), "--ok", "3", "--ot", "2",
           "--steps", "6"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("MGPU ")][-1]
    r = json.loads(line[5:])
    assert r["steps"] == ref["steps"]
    assert r["ndofs_h1_global"] == ref["ndofs_h1_global"] and r["ne_global"] == ref["ne_global"]
    assert len(r["hist"]) == len(ref["hist"])
    for (ti, e), (ti0, e0) in zip(r["hist"], ref["hist"]):
        assert ti == ti0 and abs(e - e0) <= 1e-9 * abs(e0), (ti, e, e0)

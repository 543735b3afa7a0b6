#!/usr/bin/env python
"""Every kernel of the product once on tiny meshes, for compute-sanitizer (memcheck / racecheck / initcheck):

    compute-sanitizer --tool racecheck python tools/sanitize_smoke.py [--quick]

Covers the default launch variants of every order (ok 1..5, 2D and 3D), the lagb_tune_set variants of the tuned
3D kernels (mass keys 0 / 3, Force key 1, QUpdate key 2, brick paths key 6), the direct L2 solve and the CG fallback,
ComputeDensity, and the programmatic-dependent-launch PCG chain.  Prints the git commit it ran on.
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from laghos_b200.api import Problem, Context, LagbError  # noqa: E402

quick = "--quick" in sys.argv
try:
    head = subprocess.check_output(["git", "rev-parse", "--short", "HEAD"], cwd=ROOT, text=True).strip()
except Exception:
    head = os.environ.get("LAGB_HEAD", "unknown")
print("commit", head, flush=True)

CASES = [("cube01_hex", 1, 1, 3, 2), ("cube01_hex", 0, 1, 2, 1), ("square01_quad", 1, 0, 2, 1),
         ("cube01_hex", 0, 1, 4, 3), ("cube01_hex", 0, 1, 5, 4), ("cube01_hex", 0, 1, 1, 0), ("square01_quad", 1, 1, 3, 2)]
if quick:
    CASES = CASES[:3]


def run_ops(P, c, tag):
    rng = np.random.default_rng(0)
    S = P.S0.copy()
    nv = P.h1_vsize
    S[nv:2 * nv] = 0.01 * rng.uniform(-1, 1, nv)
    S[2 * nv:] = rng.uniform(0.5, 1.5, P.ndofs_l2)
    dt = c.qupdate(c.dev(S))
    v = c.dev(rng.uniform(-1, 1, nv))
    e = c.dev(rng.uniform(0.5, 1.5, P.ndofs_l2))
    f = c.force_mult(e)
    ft = c.force_mult_transpose(v)
    c.vmass_mult(c.dev(rng.uniform(-1, 1, P.ndofs_h1)), 0)
    if P.dim == 3:
        c.vmass_mult_all(v)
    c.emass_mult(e)
    x, its = c.pcg_vmass_all(v)
    xe, it2 = c.cg_emass(e)
    c.compute_density(c.dev(S[:nv]))
    c.sync()
    print(tag, "dt", dt, "its", its, it2, float(f.abs().sum()), float(ft.abs().sum()), flush=True)


for (mesh, rs, problem, ok, ot) in CASES:
    P = Problem(mesh, rs, problem, ok, ot)
    c = Context(P)
    run_ops(P, c, f"{mesh} ok{ok} default")
    if P.dim == 3 and not quick:
        for key, vals in ((0, (1, 2, 3, 4, 5)), (3, (1, 2, 3, 4)), (1, (1, 2, 3, 4, 5)), (2, (1, 2, 3, 4, 5)), (6, (2, 3, 4)),
                          (10, (1,)), (12, (1,))):
            for val in vals:
                c.tune(key, val)
                try:
                    run_ops(P, c, f"{mesh} ok{ok} tune {key}={val}")
                except LagbError as ex:       # a variant that does not exist / fit at this order
                    print(f"{mesh} ok{ok} tune {key}={val}: unavailable ({ex})", flush=True)
                c.tune(key, 0)
    c.close()
print("sanitize smoke done")

#!/usr/bin/env python
"""Every tuned kernel once on a tiny mesh, for compute-sanitizer (memcheck / racecheck / initcheck):

    compute-sanitizer --tool racecheck python tools/sanitize_smoke.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from laghos_b200.api import Problem, Context  # noqa: E402

for (mesh, rs, problem, ok, ot) in [("cube01_hex", 1, 1, 3, 2), ("cube01_hex", 0, 1, 2, 1), ("square01_quad", 1, 0, 2, 1)]:
    P = Problem(mesh, rs, problem, ok, ot)
    c = Context(P)
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
    m1 = c.vmass_mult(c.dev(rng.uniform(-1, 1, P.ndofs_h1)), 0)
    if P.dim == 3:
        m3 = c.vmass_mult_all(v)
    em = c.emass_mult(e)
    x, its = c.pcg_vmass_all(v)
    xe, it2 = c.cg_emass(e)
    c.sync()
    print(mesh, ok, "dt", dt, "its", its, it2, float(f.abs().sum()), float(ft.abs().sum()), flush=True)
    c.close()
print("sanitize smoke done")

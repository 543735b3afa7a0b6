import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pyoracle
from laghos_b200.api import Problem, Context
rs = int(sys.argv[1]) if len(sys.argv) > 1 else 4
P = Problem("cube01_hex", rs, 1, 3, 2)
O = pyoracle.Oracle("cube01_hex", rs, 1, 3, 2, nthreads=os.cpu_count())
c = Context(P)
nv, nl = P.h1_vsize, P.ndofs_l2
n1 = int(P.info.nelem[0])
def rel(a, b): return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))
for name, xa, va, seed in [("ic", 0, 0, 1), ("x", 0.1, 0, 1), ("v1", 0, 1.0, 1), ("v1e-2", 0, 1e-2, 1), ("xv", 0.1, 1.0, 4), ("xv2", 0.1, 1.0, 5), ("xv3", 0.1, 1.0, 6)]:
    rng = np.random.default_rng(seed)
    S = P.S0.copy()
    S[:nv] += xa * (0.5 / (n1 * 3)) * rng.uniform(-1, 1, nv)
    S[nv:2 * nv] = va * rng.uniform(-1, 1, nv)
    if va or xa:
        S[2 * nv:] = rng.uniform(0.5, 1.5, nl)
    dr = O.qupdate(S); dg = c.qupdate(c.dev(S))
    print(name, dr, dg, abs(dr - dg) / dr, rel(c.qdata(0).cpu().numpy(), O.qdata(0)), flush=True)

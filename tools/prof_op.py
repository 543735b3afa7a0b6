#!/usr/bin/env python
"""Run ONE operator a few times (for `ncu -k regex:...` captures and quick timings).

    python tools/prof_op.py --op mass3 [--tune 4=2 --tune 6=1] [--reps 3] [--rs 5] [--ok 3]

ops: mass3 (lagb_vmass_mult_all), mass1 (lagb_vmass_mult), force, forcet, q (lagb_qupdate_async),
l2 (lagb_emass_mult), pcg (lagb_pcg_vmass_all), cgl2 (lagb_cg_emass).  Prints the median CUDA-event time.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--op", required=True)
    ap.add_argument("--tune", action="append", default=[], help="key=value for lagb_tune_set")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--rs", type=int, default=5)
    ap.add_argument("--ok", type=int, default=3)
    ap.add_argument("--mesh", default="cube01_hex")
    ap.add_argument("--problem", type=int, default=1)
    ap.add_argument("--sedov-steps", type=int, default=0,
                    help="take the state after this many RK4 steps of the real run instead of the perturbed IC")
    args = ap.parse_args()
    import numpy as np
    import torch
    from laghos_b200.api import Problem, Context
    P = Problem(args.mesh, args.rs, args.problem, args.ok, args.ok - 1)
    c = Context(P)
    for kv in args.tune:
        k, v = kv.split("=")
        c.tune(int(k), int(v))
    nd, nl, nv = P.ndofs_h1, P.ndofs_l2, P.h1_vsize
    rng = np.random.default_rng(1)
    S = P.S0.copy()
    S[nv:2 * nv] = 0.01 * rng.uniform(-1, 1, nv)
    S[2 * nv:] = rng.uniform(0.5, 1.5, nl)
    if args.sedov_steps > 0:
        from laghos_b200.api import run
        S = run(mesh=args.mesh, rs=args.rs, problem=args.problem, ok=args.ok, ot=args.ok - 1,
                max_tsteps=args.sedov_steps, t_final=1e9, want_state=True)["S"]
    dS = c.dev(S)
    v = c.dev(rng.uniform(-1, 1, nv))
    e = c.dev(rng.uniform(0.5, 1.5, nl))
    x1 = c.dev(rng.uniform(-1, 1, nd))
    yv, ye, y1 = c.empty(nv), c.empty(nl), c.empty(nd)
    c.qupdate(dS)
    xs = c.zeros(nv)

    def pcg():
        xs.zero_()
        c.pcg_vmass_all(v, xs)

    ops = {
        "mass3": lambda: c.lib.lagb_vmass_mult_all(c.h, c._p(v), c._p(yv)),
        "mass1": lambda: c.lib.lagb_vmass_mult(c.h, -1, c._p(x1), c._p(y1)),
        "force": lambda: c.lib.lagb_force_mult(c.h, c._p(e), c._p(yv)),
        "forcet": lambda: c.lib.lagb_force_mult_transpose(c.h, c._p(v), c._p(ye)),
        "q": lambda: c.lib.lagb_qupdate_async(c.h, c._p(dS), 0.5),
        "l2": lambda: c.lib.lagb_emass_mult(c.h, c._p(e), c._p(ye)),
        "pcg": pcg,
        "cgl2": lambda: c.cg_emass(e),
    }
    fn = ops[args.op]
    fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.reps + 1)]
    ev[0].record()
    for i in range(args.reps):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(args.reps))
    print(f"{args.op} tune={args.tune}: median {1e3 * ts[len(ts) // 2]:.1f} us over {args.reps} reps", flush=True)
    c.close()


if __name__ == "__main__":
    main()

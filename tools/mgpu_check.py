#!/usr/bin/env python
"""torchrun helper: element-partitioned run over WORLD_SIZE GPUs, rank 0 prints one JSON line.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        tools/mgpu_check.py --pgrid 2,1,1 --mesh cube01_hex --rs 2 --problem 1 --ok 3 --ot 2 --steps 6
"""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pgrid", default="2,1,1")
    ap.add_argument("--mesh", default="cube01_hex")
    ap.add_argument("--rs", type=int, default=2)
    ap.add_argument("--problem", type=int, default=1)
    ap.add_argument("--ok", type=int, default=3)
    ap.add_argument("--ot", type=int, default=2)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--cg-tol", type=float, default=1e-12)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from laghos_b200 import load_library
    from laghos_b200.api import run
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = load_library()
    idt = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        buf = ctypes.create_string_buffer(128)
        assert lib.lagb_nccl_unique_id(buf) == 0, lib.lagb_last_error()
        idt = torch.tensor(list(buf.raw), dtype=torch.uint8)
    idt = idt.cuda()
    dist.broadcast(idt, 0)
    pg = tuple(int(v) for v in args.pgrid.split(","))
    r = run(mesh=args.mesh, rs=args.rs, problem=args.problem, ok=args.ok, ot=args.ot, max_tsteps=args.steps,
            t_final=1e9, cg_tol=args.cg_tol, device=local, rank=rank, nranks=world, pgrid=pg,
            nccl_id=bytes(idt.cpu().tolist()), hist_cap=1024)
    if rank == 0:
        print("MGPU " + json.dumps(dict(steps=r["steps"], e_norm=r["e_norm"], dt=r["dt"], hist=r["hist"],
                                        ndofs_h1_global=r["ndofs_h1_global"], ne_global=r["ne_global"],
                                        H1iter=r["H1iter"])), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

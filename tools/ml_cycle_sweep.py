"""iterations and solve time of CG + FS_PC_MLRBM for different cycle indices (fs_set_ml_options gamma: 2 = W, 1 = V,
21 = W on the first lattice / V below, ...) and dense thresholds, one GPU or torchrun (one strip per rank)
    python tools/ml_cycle_sweep.py [nodes] [q|t]"""
import json, os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import numpy as np
import torch
import fem_shell_b200 as fsb

nodes = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
kind = sys.argv[2] if len(sys.argv) > 2 else "q"
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
nid = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    ids = [fsb.FemShell.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    nid = ids[0]
m = fsb.meshgen(kind, nodes - 1, nodes * world - 1, 0.0, 0.0, 10.0, 10.0 * world, (1, 1, 1, 1), 300.0, 2, 1)
s = fsb.FemShell(device=lr, rank=rank, world=world, nccl_id=nid)
s.set_material(0.3, 1.0e7, 0.5)
s.set_mesh(m["xyz"], m["etype"], m["eptr"], m["enodes"], m["bc"])
s.set_nodal_loads(m["forces"])
s.assemble()
s.build_rhs(1.0)
rows = []
for gamma in [int(g) for g in os.environ.get("FS_SWEEP_GAMMAS", "2,21,22,31,1,32").split(",")]:
    s.set_ml_options(gamma=gamma)
    s.solve(rtol=1e-8, max_its=5000, pc=fsb.PC_MLRBM, warm_start=False, allow_not_converged=True)   # set-up + capture
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    info = s.solve(rtol=1e-8, max_its=5000, pc=fsb.PC_MLRBM, warm_start=False, allow_not_converged=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    rows.append({"gamma": gamma, "iterations": info.iterations, "solve_ms": info.solve_ms, "wall_s": dt, "ms_per_iteration": info.solve_ms / max(1, info.iterations),
                 "converged": info.status == 0, "setup_ms": s.ml_info()["setup_ms"]})
    if rank == 0:
        print(json.dumps(rows[-1]), flush=True)
if rank == 0:
    print(json.dumps({"world": world, "nodes_per_rank": nodes, "kind": kind, "cells": s.ml_info()["cells"], "rows": rows}))
s.close()

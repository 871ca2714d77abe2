"""stage times of one multilevel-preconditioned CG iteration (FS_ML_PROFILE=1: eager iterations, CUDA events between
the stages), on one GPU or under torchrun on N (weak scaling: one NODES x NODES strip per rank, like bench.py)
    python tools/ml_stage_profile.py [nodes] [q|t] [nodes_x]
    python -m torch.distributed.run --nproc-per-node N ... tools/ml_stage_profile.py [nodes] [q|t] [nodes_x]
nodes = node rows per rank, nodes_x = nodes per row (default: nodes); c3 on 8 GPUs: 500 t 4000"""
import json, os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
os.environ["FS_ML_PROFILE"] = "1"
import numpy as np
import torch
import fem_shell_b200 as fsb

nodes = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
kind = sys.argv[2] if len(sys.argv) > 2 else "q"
nodes_x = int(sys.argv[3]) if len(sys.argv) > 3 else nodes
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
nid = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    ids = [fsb.FemShell.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    nid = ids[0]
m = fsb.meshgen(kind, nodes_x - 1, nodes * world - 1, 0.0, 0.0, 10.0, 10.0 * (nodes * world - 1) / (nodes_x - 1), (1, 1, 1, 1), 300.0, 2, 1)
s = fsb.FemShell(device=lr, rank=rank, world=world, nccl_id=nid)
s.set_material(0.3, 1.0e7, 0.5)
s.set_mesh(m["xyz"], m["etype"], m["eptr"], m["enodes"], m["bc"])
s.set_nodal_loads(m["forces"])
s.assemble()
s.build_rhs(1.0)
t0 = time.perf_counter()
info = s.solve(rtol=1e-8, max_its=5000, pc=fsb.PC_MLRBM, warm_start=False)
wall = time.perf_counter() - t0
p = s.ml_profile()
if rank == 0:
    tot = sum(v for k, v in p.items() if k not in ("iterations", "lattice_level2_visit1", "lattice_level2_visit2"))
    print(json.dumps({"world": world, "nodes_per_rank": [nodes_x, nodes], "kind": kind, "iterations": info.iterations, "solve_wall_s": wall,
                      "ml": s.ml_info(), "stage_ms_per_iteration": p, "sum_ms": tot, "comm": "peer" if s.comm_mode() == fsb.COMM_PEER else "nccl"}))
s.close()

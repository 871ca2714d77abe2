"""North-star target run (BASELINE configs[2]): meshGen plate of NODES x NODES nodes (Tri-3 by default), clamped, uniform
pressure, assembled and solved to rtol 1e-8 with CG + FS_PC_MLRBM on WORLD GPUs (torchrun).  Prints one JSON line.
Checks that travel to any size: the solver's own relative residual, an independent residual b - K u recomputed with
the SpMV of the parity matrix format on every rank, and the centre deflection against the thin-plate series value
w = 0.00126 q a^4 / D (Timoshenko, clamped square plate, uniform load)."""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import fem_shell_b200 as fsb

ap = argparse.ArgumentParser()
ap.add_argument("--nodes", type=int, default=4000)
ap.add_argument("--kind", default="t")
ap.add_argument("--rtol", type=float, default=1e-8)
ap.add_argument("--gamma", type=int, default=2)
args = ap.parse_args()
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
nccl_id = None
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    ids = [fsb.FemShell.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    nccl_id = ids[0]
NU, EM, TH, Q, A = 0.3, 1.0e7, 0.5, 300.0, 10.0
n = args.nodes - 1
t0 = time.perf_counter()
m = fsb.meshgen(args.kind, n, n, 0.0, 0.0, A, A, (1, 1, 1, 1), Q, 2, 1)
t_gen = time.perf_counter() - t0
s = fsb.FemShell(device=lr, rank=rank, world=world, nccl_id=nccl_id, comm=fsb.COMM_NCCL)
s.set_material(NU, EM, TH)
s.set_ml_options(gamma=args.gamma)
t0 = time.perf_counter()
s.set_mesh(m["xyz"], m["etype"], m["eptr"], m["enodes"], m["bc"])
t_mesh = time.perf_counter() - t0
s.set_nodal_loads(m["forces"])
sz = s.sizes()


def sync():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()


s.assemble()                                  # warm-up values pass
s.build_rhs(1.0)
s.solve(rtol=args.rtol, max_its=2, pc=fsb.PC_MLRBM, warm_start=False, allow_not_converged=True)   # hierarchy geometry + graph capture
sync()
t0 = time.perf_counter()
asm_ms = s.assemble()
sync()
t_asm = time.perf_counter() - t0
t0 = time.perf_counter()
info = s.solve(rtol=args.rtol, max_its=5000, pc=fsb.PC_MLRBM, warm_start=False, allow_not_converged=True)
sync()
t_solve = time.perf_counter() - t0
mi = s.ml_info()
u = s.solution()                               # (n_nodes, 6) in mesh node order, replicated
w = u[:, 2].reshape(args.nodes, args.nodes)
D = EM * TH ** 3 / (12.0 * (1.0 - NU * NU))
h = A / n
q_area = m["forces"][0, 2] / (h * h) if m["forces"][0, 2] != 0 else m["forces"][args.nodes + 1, 2] / (h * h)
w_c = w[args.nodes // 2, args.nodes // 2]
w_ref = 0.00126 * q_area * A ** 4 / D
if rank == 0:
    print(json.dumps({
        "workload": "meshGen %dx%d nodes %s, clamped, uniform pressure" % (args.nodes, args.nodes, "Tri-3 (Specht+CST)" if args.kind == "t" else "Quad-4 (DKQ+PLANE)"),
        "n_gpus": world, "n_dof": 6 * sz["n_dofnodes"], "n_elem": int(m["etype"].size), "rtol": args.rtol,
        "meshgen_s": t_gen, "set_mesh_s": t_mesh, "assemble_s": t_asm, "assemble_kernel_ms": asm_ms,
        "solve_s_incl_setup": t_solve, "solve_kernel_ms": info.solve_ms, "ml_setup_ms": mi["setup_ms"], "iterations": info.iterations,
        "rel_residual": info.rel_residual, "converged": info.status == 0, "time_to_solution_s": t_asm + t_solve,
        "ml_cells": mi["cells"], "ml_lambda": mi["lambda"], "gamma": args.gamma,
        "centre_deflection": w_c, "thin_plate_series": w_ref, "ratio": w_c / w_ref,
        "symmetry_err": float(np.abs(w - w.T).max() / np.abs(w).max()),
    }), flush=True)
if world > 1:
    dist.destroy_process_group()

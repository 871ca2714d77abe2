#!/usr/bin/env python
"""bench.py -- headline benchmark of the fem-shell hot path on B200 (contract: see DESIGN.md section 6).

    python bench.py --gpus 1 --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --steps K --warmup W    # CPU restatement of the reference (oracle)
    torchrun ... bench.py --gpus N ...                       # one rank per GPU: weak-scaled c2 strips (value) + the
                                                             # 96 M-DOF Tri-3 plate of BASELINE configs[2] (metrics.c3)

Workload at N=1 = BASELINE.json configs[1]: meshGen square plate, 1000 x 1000 nodes of DKQ+PLANE
Quad-4 (998 001 elements, 6 000 000 DOF), clamped edges, uniform pressure, E=1e7 nu=0.3 t=0.5.
At N>1 every rank owns one such 1000 x 1000-node strip of a 1000 x (1000 N) plate (weak scaling).

A step = one pass of the solve hot path over one load case: rhs for a new pressure amplitude
(1+sin(tau/25.01), fluid_solver.cpp:192) followed by exactly --iters Jacobi-PCG iterations on the
device-resident stiffness matrix (2.6 GB, far larger than the 126 MB L2, so nothing is cache
resident between steps).  value = DOF x iterations / second over all ranks.  The same JSON line
also carries elements assembled/s (values pass timed over K repetitions), the SpMV roofline, the
end-to-end number through the host-buffer plugin call, a CPU baseline and time-to-solution.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NU, EM, THICK, QLOAD, PLATE = 0.3, 1.0e7, 0.5, 300.0, 10.0
METRIC = "CG DOF-iterations/s"
UNIT = "DOF-iterations/s"


def claim_stdout():
    """stdout carries exactly ONE JSON line: from here on file descriptor 1 is stderr for everything else that writes to
    it (NCCL prints its version banner to stdout when the box sets NCCL_DEBUG); returns the descriptor of the real stdout"""
    sys.stdout.flush()
    fd = os.dup(1)
    os.dup2(2, 1)
    return fd


def emit_line(fd, line):
    sys.stdout.flush()
    os.write(fd, (json.dumps(line) + "\n").encode())


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel, override):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture
    (profiles/ncu_traffic.json; same c2 per-GPU workload), or the --traffic override"""
    if override is not None:
        return override
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        return float(json.load(open(p))[kernel]["traffic"])
    except (OSError, KeyError, ValueError):
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region"""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def nvlink_kib(index):
    """NVLink data counters of one GPU, summed over its links: (tx KiB, rx KiB) or None.  Read before and after the timed
    region: the bytes the peer-window kernels (or NCCL) really moved, from the link side.  NVML field values first, the
    text of `nvidia-smi nvlink -gt d` as a fallback."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        v = pynvml.nvmlDeviceGetFieldValues(h, [(pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX, 0xFFFFFFFF),
                                                 (pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_RX, 0xFFFFFFFF)])
        if all(x.nvmlReturn == 0 for x in v):
            return int(v[0].value.ullVal), int(v[1].value.ullVal)
    except Exception:
        pass
    try:
        out = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", str(index)], capture_output=True, text=True, timeout=20).stdout
        tx = rx = 0
        seen = False
        for ln in out.splitlines():
            ln = ln.strip()
            if "Data Tx:" in ln:
                tx += int(ln.split("Data Tx:")[1].split()[0]); seen = True
            elif "Data Rx:" in ln:
                rx += int(ln.split("Data Rx:")[1].split()[0]); seen = True
        return (tx, rx) if seen else None
    except Exception:
        return None


def make_workload(gen, nodes_x, nodes_y, kind="q"):
    """meshGen: q nx ny 0 0 Lx Ly 1,1,1,1 q0 2 1 z  (clamped = boundary id 1, uniform load).  gen = the product's
    host generator (fsb.meshgen) on our arm, the oracle's numpy restatement (fso.meshgen) on the reference arm."""
    return gen(kind, nodes_x - 1, nodes_y - 1, 0.0, 0.0, PLATE, PLATE * (nodes_y - 1) / (nodes_x - 1),
               (1, 1, 1, 1), QLOAD, 2, 1)


def workload_config(nx, ny, world, iters):
    """the `config` object of the JSON line -- built by ONE function for both arms, so that the reference arm runs
    (and says it runs) exactly the workload of ours"""
    n_elem, n_dof = (nx - 1) * (ny - 1), 6 * nx * ny
    return {"workload": "BASELINE configs[1]: meshGen %dx%d nodes DKQ+PLANE Quad-4 (%d elements, %d DOF), clamped, uniform pressure%s"
                        % (nx, ny, n_elem, n_dof, "" if world == 1 else " = one %dx%d-node strip per GPU" % (nx, nx)),
            "iters_per_step": iters, "pc": "jacobi", "dof_order": "first_encounter",
            "l2_policy": "inputs larger than L2 (the stiffness matrix, > 1 GB per GPU, is streamed from memory every iteration)",
            "parallelism": "node-block strips x%d" % world}


def host_threads():
    """the cores this process may use -- NOT omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


TTS_SMALL_NODES = 150


def tts_small_reference(fso, threads):
    """bounded time-to-solution both arms run: assemble + Jacobi-PCG to rtol 1e-8 (the reference's documented
    -ksp_type cg -pc_type jacobi, fs.cpp:138) on a 150x150-node Quad-4 plate the CPU finishes in seconds"""
    n = TTS_SMALL_NODES
    mesh, forces = make_workload(fso.meshgen, n, n)
    t0 = time.perf_counter()
    sysm = fso.assemble(mesh, forces, NU, EM, THICK, threads=threads)
    t_asm = time.perf_counter() - t0
    t0 = time.perf_counter()
    _, its, rel = fso.pcg(sysm, rtol=1e-8, max_its=2000000, threads=threads)
    t_cg = time.perf_counter() - t0
    return {"workload": "meshGen %dx%d nodes Quad-4 (%d DOF), clamped, uniform pressure: assemble + Jacobi-PCG to rtol 1e-8" % (n, n, 6 * n * n),
            "seconds": t_asm + t_cg, "assemble_s": t_asm, "solve_s": t_cg, "iterations": its, "rel_residual": rel, "threads": threads}


def spmv_bytes(n_dof, n_blocks, matrix_bytes):
    """bytes one SpMV launch must move: the matrix in the format actually streamed (fs_get_spmv_format:
    values + column ids + row/slice pointers) + x read once + y written once; and the SURVEY.md section 8d
    figure for the same matrix as a scalar CSR with 32-bit column ids (explicit zeros of the 6x6 blocks included)"""
    actual = matrix_bytes + 8 * n_dof + 8 * n_dof
    csr_equiv = 12 * 36 * n_blocks + 20 * n_dof
    return actual, csr_equiv


# ------------------------------------------------------------------------------------------------
def run_reference(args):
    """the reference's CPU implementation of the path, restated (oracle/fs_oracle.c; the reference itself needs
    libMesh+PETSc+MPI, which do not exist in this image) on ALL the host cores of the box, on our arm's config.
    Nothing of the product is loaded in this process: the mesh comes from the oracle's own meshGen restatement.
    A step = rhs for a new pressure amplitude + Jacobi-PCG iterations on the assembled system, like ours; when the
    whole --steps/--warmup run would not end within --ref-budget-s, a step runs the first k of the config's
    iterations (a bounded sample; the metric is a rate and the per-call set-up is reported beside it)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import fso
    world = max(1, args.gpus)
    threads = args.ref_threads or host_threads()
    nx, ny = args.ref_nodes, args.ref_nodes * world
    mesh, forces = make_workload(fso.meshgen, nx, ny)
    t0 = time.perf_counter()
    sysm = fso.assemble(mesh, forces, NU, EM, THICK, threads=threads)
    t_asm = time.perf_counter() - t0
    n_dof = 6 * sysm.n_dofnodes
    # calibration: per-call set-up (diagonal, r0 = b - A x0) and per-iteration cost
    t0 = time.perf_counter(); fso.pcg(sysm, rtol=1e-30, max_its=0, threads=threads); t_setup = time.perf_counter() - t0
    t0 = time.perf_counter(); fso.pcg(sysm, rtol=1e-30, max_its=4, threads=threads); t_it = max(1e-9, (time.perf_counter() - t0 - t_setup) / 4)
    budget_per_step = args.ref_budget_s / max(1, args.steps + args.warmup)
    iters = args.ref_iters if args.ref_iters > 0 else int(max(5, min(args.iters, (budget_per_step - t_setup) / t_it)))
    for k in range(args.warmup):
        fso.pcg(sysm, b=sysm.rhs * (1.0 + np.sin(k / 25.01)), rtol=1e-30, max_its=iters, threads=threads)
    done = 0
    t0 = time.perf_counter()
    for k in range(args.steps):
        _, its, _ = fso.pcg(sysm, b=sysm.rhs * (1.0 + np.sin(k / 25.01)), rtol=1e-30, max_its=iters, threads=threads)
        done += its
    dt = time.perf_counter() - t0
    assert done == iters * args.steps, "the oracle's PCG stopped early (%d of %d iterations)" % (done, iters * args.steps)
    value = n_dof * done / dt
    sample = ("the config's %dx%d-node system (%d DOF): %d of its %d Jacobi-PCG iterations per step (OpenMP over rows, %d threads); "
              "per-call set-up %.3f s, %.4f s per iteration" % (nx, ny, n_dof, iters, args.iters, threads, t_setup, t_it))
    tts_small = None if args.tts == "off" else tts_small_reference(fso, threads)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(nx, ny, world, args.iters),
        "metrics": {"cg_dof_iterations_per_s": value, "elements_assembled_per_s": mesh.n_elem / t_asm, "assemble_ms": 1e3 * t_asm,
                    "time_to_solution_bounded": tts_small, "iters_run_per_step": iters},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "elements_per_s": mesh.n_elem / t_asm},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
def dist_parity_preflight(fsb, torch, dist, rank, world, local_rank, comm):
    """world-N correctness the single-GPU test box cannot show: a 97x121-node Tri-3 plate assembled and solved on all
    ranks (block-Jacobi PCG and the multilevel PCG), compared with the oracle's direct solve of the oracle's matrix.
    Returns max rel-L2 displacement error over the two solves (rank 0: also the iteration counts)."""
    from oracle import fso
    mesh, forces = fso.meshgen("t", 96, 120, 0.0, 0.0, PLATE, PLATE * 1.25, (1, 1, 1, 1), QLOAD, 2, 1)
    ref = fso.assemble(mesh, forces, NU, EM, THICK)
    uo = fso.direct_solve(mesh, ref)
    ids = [fsb.FemShell.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    s = fsb.FemShell(device=local_rank, rank=rank, world=world, nccl_id=ids[0], comm=comm)
    s.set_material(NU, EM, THICK)
    s.set_mesh(mesh.xyz, mesh.etype, mesh.eptr, mesh.enodes, mesh.bc)
    s.set_nodal_loads(forces)
    s.assemble()
    sz = s.sizes()
    ob, oe = sz["own_begin"], sz["own_end"]
    rowptr, colidx, vals = s.export_csr()
    rr, rc, rv = ref.csr()
    lo, hi = rr[6 * ob], rr[6 * oe]
    csr_ok = bool(np.array_equal(rowptr, rr[6 * ob:6 * oe + 1] - lo) and np.array_equal(colidx, rc[lo:hi])
                  and np.abs(vals - rv[lo:hi]).max() <= 1e-12 * np.abs(rv).max())
    out = {"csr_rows_match_oracle": csr_ok}
    worst = 0.0
    for name, pc in (("bjacobi6", fsb.PC_BJACOBI6), ("mlrbm", fsb.PC_MLRBM)):
        info = s.solve(rtol=1e-12, max_its=400000, pc=pc, warm_start=False)
        err = float(np.linalg.norm(s.solution() - uo) / np.linalg.norm(uo))
        out[name] = {"iterations": info.iterations, "rel_l2_vs_oracle": err}
        worst = max(worst, err)
    s.close()
    t = torch.tensor([worst, 0.0 if csr_ok else 1.0], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out["rel_l2"] = float(t[0].item())
    out["csr_rows_match_oracle"] = bool(t[1].item() == 0.0)
    out["workload"] = "meshGen 97x121 nodes Tri-3, clamped, uniform pressure, world %d" % world
    return out


def run_c3(args, fsb, torch, dist, rank, world, local_rank, comm, peak):
    """BASELINE configs[2]: the 4000x4000-node DKT(Specht)+CST plate (96 M DOF) partitioned over the N GPUs of the box
    (strong scaling: the plate is fixed, N varies), assembled and solved to rtol 1e-8 (CG + FS_PC_MLRBM)."""
    nodes = args.c3_nodes

    def sync():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    t0 = time.perf_counter()
    m = make_workload(fsb.meshgen, nodes, nodes, kind="t")
    t_gen = time.perf_counter() - t0
    ids = [fsb.FemShell.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    s = fsb.FemShell(device=local_rank, rank=rank, world=world, nccl_id=ids[0], comm=comm)
    s.set_material(NU, EM, THICK)
    sync()
    t_first0 = time.perf_counter()
    t0 = time.perf_counter()
    s.set_mesh(m["xyz"], m["etype"], m["eptr"], m["enodes"], m["bc"])
    sync()
    t_mesh = time.perf_counter() - t0
    s.set_nodal_loads(m["forces"])
    sz = s.sizes()
    n_elem = int(m["etype"].size)
    t0 = time.perf_counter()
    s.assemble()
    sync()
    t_first_asm = time.perf_counter() - t0
    info = s.solve(rtol=1e-8, max_its=5000, pc=fsb.PC_MLRBM, warm_start=False, allow_not_converged=True)
    sync()
    t_first = time.perf_counter() - t_first0          # cold: mesh set-up, schedules, hierarchy, graph capture, solve
    # warm repetition = the reference's equation_systems.solve() (fs.cpp:138): zero K, assemble, solve
    sync()
    t0 = time.perf_counter()
    asm_ms = s.assemble()
    sync()
    t_asm = time.perf_counter() - t0
    t0 = time.perf_counter()
    info = s.solve(rtol=1e-8, max_its=5000, pc=fsb.PC_MLRBM, warm_start=False, allow_not_converged=True)
    sync()
    t_solve = time.perf_counter() - t0
    mi = s.ml_info()
    spmv_ms = s.bench_spmv(20)
    fmt = s.spmv_format()
    n_own = sz["own_end"] - sz["own_begin"]
    spmv_b = fmt["matrix_bytes"] + 16 * 6 * n_own
    t = torch.tensor([spmv_ms, asm_ms], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    spmv_ms, asm_ms = float(t[0].item()), float(t[1].item())
    u = s.solution()
    w = u[:, 2].reshape(nodes, nodes)
    D = EM * THICK ** 3 / (12.0 * (1.0 - NU * NU))
    h = PLATE / (nodes - 1)
    w_ref = 0.00126 * (m["forces"][nodes + 1, 2] / (h * h)) * PLATE ** 4 / D     # Timoshenko, clamped square plate
    out = {
        "workload": "BASELINE configs[2]: meshGen %dx%d nodes DKT(Specht)+CST Tri-3 (%d elements, %d DOF), clamped, uniform pressure, "
                    "node-block strips over %d GPUs (strong scaling)" % (nodes, nodes, n_elem, 6 * sz["n_dofnodes"], world),
        "n_gpus": world, "n_dof": 6 * sz["n_dofnodes"], "n_elem": n_elem, "rtol": 1e-8, "pc": "mlrbm",
        "meshgen_s": t_gen, "set_mesh_s": t_mesh, "first_assemble_s": t_first_asm, "time_to_first_solution_s": t_first,
        "assemble_ms": asm_ms, "elements_assembled_per_s": n_elem / (asm_ms * 1e-3),
        "assemble_s": t_asm, "solve_s": t_solve, "ml_setup_ms": mi["setup_ms"], "iterations": info.iterations,
        "rel_residual": info.rel_residual, "converged": info.status == 0,
        "time_to_solution_s": t_asm + t_solve,
        "spmv_ms_per_gpu": spmv_ms, "spmv_gbs_per_gpu": spmv_b / (spmv_ms * 1e-3) / 1e9, "spmv_frac_of_hbm_peak": spmv_b / (spmv_ms * 1e-3) / 1e9 / peak,
        "spmv_nz_per_block": fmt["nz_per_block"], "ml_cells": mi["cells"], "ml_distributed_levels": mi["distributed_levels"],
        "centre_deflection": float(w[nodes // 2, nodes // 2]), "thin_plate_series": w_ref, "deflection_ratio": float(w[nodes // 2, nodes // 2] / w_ref),
        "symmetry_err": float(np.abs(w - w.T).max() / np.abs(w).max()),
        "comm": "peer" if s.comm_mode() == fsb.COMM_PEER else "nccl",
    }
    s.close()
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    import fem_shell_b200 as fsb

    json_fd = claim_stdout()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    torch.cuda.set_device(local_rank)
    nccl_id = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        ids = [fsb.FemShell.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        nccl_id = ids[0]

    nx = args.nodes
    ny = args.nodes * world                      # weak scaling: one 1000 x 1000-node strip per GPU
    m = make_workload(fsb.meshgen, nx, ny)
    n_nodes, n_elem = m["xyz"].shape[0], m["etype"].size
    comm = {"auto": fsb.COMM_AUTO, "nccl": fsb.COMM_NCCL, "peer": fsb.COMM_PEER}[args.comm]

    def new_context():
        c = fsb.FemShell(device=local_rank, rank=rank, world=world, nccl_id=nccl_id if world == 1 else fresh_id(), comm=comm)
        c.set_material(NU, EM, THICK)
        c.set_assembly_mode(fsb.ASM_GATHER if args.asm == "gather" else fsb.ASM_COLORED)
        c.set_spmv_format(fsb.SPMV_FULL if args.spmv == "full" else fsb.SPMV_AUTO)
        return c

    def fresh_id():                              # an ncclUniqueId serves one communicator
        ids = [fsb.FemShell.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        return ids[0]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- time to the FIRST solution of a new mesh (cold context): mesh ingestion (DOF map, partition, pattern,
    # colouring, upload), gather schedule, first values pass, multilevel hierarchy, graph capture, solve to 1e-8,
    # displacements on the host.  The context then serves the rest of the run. ----
    s = new_context()
    barrier()
    t_first0 = time.perf_counter()
    t0 = time.perf_counter()
    s.set_mesh(m["xyz"], m["etype"], m["eptr"], m["enodes"], m["bc"])
    t_setup = time.perf_counter() - t0
    s.set_nodal_loads(m["forces"])
    t0 = time.perf_counter()
    s.assemble()
    t_first_assemble = time.perf_counter() - t0
    first = None
    if args.tts != "off":
        info = s.solve(rtol=1e-8, max_its=5000, pc=fsb.PC_MLRBM, warm_start=False, allow_not_converged=True)
        s.solution_owned(with_ids=False)
        barrier()
        first = {"seconds": time.perf_counter() - t_first0, "set_mesh_s": t_setup, "first_assemble_s": t_first_assemble,
                 "iterations": info.iterations, "converged": info.status == 0, "pc": "mlrbm", "rtol": 1e-8,
                 "includes": "fs_set_mesh, loads, first fs_assemble (gather schedule), multilevel set-up, graph capture, PCG to 1e-8, own rows D2H"}
    sz = s.sizes()
    n_dof = 6 * sz["n_dofnodes"]
    n_own = sz["own_end"] - sz["own_begin"]
    stream = torch.cuda.ExternalStream(s.stream, device=torch.device("cuda", local_rank))

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn, reps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for k in range(reps):
            fn(k)
        e1.record(stream)
        barrier()
        return max_over_ranks(e0.elapsed_time(e1))

    # ---- assembly: values pass over all elements (pattern + schedule were built once above) ----
    for _ in range(max(args.warmup, 3)):
        s.assemble()
    asm_ms = timed(lambda k: s.assemble(), args.steps) / args.steps

    # ---- the timed region: K steps of (rhs + ITERS PCG iterations) ----
    iters = args.iters
    its_done = [0]

    def check(info):
        # a solve that stopped early (breakdown, communication failure, NaN) turns every later kernel into a no-op:
        # it must not be counted as `iters` iterations
        if info.iterations != iters or info.status not in (fsb.FS_OK, fsb.FS_ERR_NOT_CONVERGED):
            raise SystemExit("timed step ran %d of %d iterations (status %d)" % (info.iterations, iters, info.status))
        its_done[0] += info.iterations

    def step(k):
        s.build_rhs(1.0 + np.sin(k / 25.01))
        check(s.solve(rtol=1e-30, max_its=iters, pc=fsb.PC_JACOBI, warm_start=False, check_every=iters, allow_not_converged=True))

    for k in range(args.warmup):
        step(k)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    its_done[0] = 0
    if world > 1:
        s.comm_stats(reset=True)
    nvl0 = nvlink_kib(local_rank) if world > 1 else None
    total_ms = timed(step, args.steps)
    nvl1 = nvlink_kib(local_rank) if world > 1 else None
    waits = s.comm_stats(reset=True) if world > 1 else None
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = total_ms / args.steps
    its_timed = its_done[0]
    value = n_dof * (its_timed / args.steps) / (ms_per_step * 1e-3)

    # ---- dominant kernel: SpMV, timed alone on the same stream right after the timed region ----
    spmv_ms = max_over_ranks(s.bench_spmv(50))
    spmv_ms_on_p = max_over_ranks(s.last_spmv_ms_on_p)
    fmt = s.spmv_format()
    comm_used = "none (single rank)" if world == 1 else ("NVLink peer windows: halo push + mailbox all-reduce inside the CG kernels"
                                                          if s.comm_mode() == fsb.COMM_PEER else "NCCL send/recv + all-reduce per iteration")
    actual_b, csr_b = spmv_bytes(6 * n_own, sz["n_blocks"], fmt["matrix_bytes"])
    spmv_kernel = "k_spmv_sell" if fmt["nz_per_block"] < 36 else "k_spmv"
    peak, peak_src = peaks()
    achieved = actual_b / (spmv_ms * 1e-3) / 1e9

    # ---- element kernel rooflines: algorithmic work per element from SURVEY.md section 8d ----
    fp64_peak = s.bench_fp64_peak()
    kflop_el, bytes_el = 15.7e3, 2630.0          # Quad-4: flops and minimum HBM bytes per element
    asm_rate = n_elem / world / (asm_ms * 1e-3)  # per GPU
    assembly_roofline = {
        "kernel": s.assembly_path(),
        "fp64": {"achieved": asm_rate * kflop_el / 1e12, "peak": fp64_peak, "unit": "TFLOP/s", "frac": asm_rate * kflop_el / 1e12 / fp64_peak,
                 "peak_source": "measured here (fs_bench_fp64_peak, dependent-FMA chains)"},
        "hbm": {"achieved": asm_rate * bytes_el / 1e9, "peak": peak, "unit": "GB/s", "frac": asm_rate * bytes_el / 1e9 / peak},
        "per_element": {"flop": kflop_el, "bytes": bytes_el},
    }

    # ---- end to end through the host-buffer plugin call (loads in, displacements out) ----
    F_host = torch.empty((n_nodes, 6), dtype=torch.float64).pin_memory()
    F_host.copy_(torch.from_numpy(m["forces"]))
    sols_host = torch.empty((n_nodes, 6), dtype=torch.float64).pin_memory()
    Fh, Sh = F_host.numpy(), sols_host.numpy()

    # N = 1: fs_solve_host returns the whole sols[6*node+var].  N > 1: every rank reads back its OWN rows
    # (fs_get_solution_owned = the distributed vector PETSc holds before build_solution_vector replicates it,
    # fs.cpp:140); replicating the full solution on all N hosts would cost N x the bytes for the same information.
    own_host = None
    if world > 1:
        own_host = torch.empty((n_own, 6), dtype=torch.float64).pin_memory()
        Oh = own_host.numpy()

    def e2e_step(k):
        if world == 1:
            check(s.solve_host(Fh, Sh, reassemble=True, rtol=1e-30, max_its=iters, pc=fsb.PC_JACOBI, warm_start=False,
                               check_every=iters, allow_not_converged=True))
        else:
            check(s.solve_host(Fh, None, reassemble=True, rtol=1e-30, max_its=iters, pc=fsb.PC_JACOBI, warm_start=False,
                               check_every=iters, allow_not_converged=True))
            s.solution_owned(out=Oh, with_ids=False)

    for k in range(min(args.warmup, 3)):
        e2e_step(k)
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        e2e_step(k)
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = n_dof * iters * args.steps / e2e_s

    # ---- time to solution (assemble + PCG to rtol 1e-8), bounded ----
    # "multilevel": CG with the smoothed-aggregation cycle FS_PC_MLRBM (values set-up inside the timed region);
    # "jacobi": the reference's documented -pc_type jacobi, capped at --tts-max-s (on for --tts-pc both|jacobi)
    # oracle_residual (N = 1): ||b - K_oracle u|| / ||b|| with the ORACLE's CSR product, and the floating-point floor
    # of that evaluation, 2 n eps || |K_oracle| |u| || / ||b||  (tests/test_gpu_fullsize.py explains the bar)
    tts = None
    orc = None
    if args.tts != "off":
        per_iter_ms = ms_per_step / iters
        cap = int(max(1000, min(args.tts_max_s * 1e3 / per_iter_ms, 5e6)))
        tts = {}
        if world == 1 and rank == 0 and not args.no_cpu:
            from oracle import fso
            th = host_threads()
            om = fso.Mesh(m["xyz"], m["etype"], m["eptr"], m["enodes"], m["bc"])
            t0 = time.perf_counter()
            osys = fso.assemble(om, m["forces"], NU, EM, THICK, threads=th)
            orc = {"fso": fso, "om": om, "sys": osys, "threads": th, "t_asm": time.perf_counter() - t0,
                   "abs": fso.System(osys.dofnode, osys.n_dofnodes, osys.mask, osys.nptr, osys.nadj, np.abs(osys.vals), osys.rhs)}

        def oracle_residual(u_nodes):
            fso, osys, th = orc["fso"], orc["sys"], orc["threads"]
            x = np.zeros(6 * osys.n_dofnodes)
            x.reshape(-1, 6)[osys.dofnode] = u_nodes
            bn = np.linalg.norm(osys.rhs)
            r = osys.rhs - fso.spmv(osys, x, threads=th)
            floor = 2 * 54 * np.finfo(np.float64).eps * np.linalg.norm(fso.spmv(orc["abs"], np.abs(x), threads=th))
            return float(np.linalg.norm(r) / bn), float(floor / bn)

        def tts_run(pc, max_its, check_every, residual=True):
            s.build_rhs(1.0)
            barrier()
            t0 = time.perf_counter()
            a_ms = s.assemble()
            info = s.solve(rtol=1e-8, max_its=max_its, pc=pc, warm_start=False, check_every=check_every, allow_not_converged=True)
            barrier()
            r = {"seconds": time.perf_counter() - t0, "assemble_ms": a_ms, "solve_ms": info.solve_ms, "iterations": info.iterations,
                 "rel_residual": info.rel_residual, "converged": info.status == 0, "rtol": 1e-8, "iteration_cap": max_its}
            if orc is not None and residual:
                r["oracle_residual"], r["oracle_residual_eval_floor"] = oracle_residual(s.solution())
            return r

        u_ml = None
        if args.tts_pc in ("ml", "both"):
            tts_run(fsb.PC_MLRBM, 4, 0, residual=False)   # the timed Jacobi steps replaced the captured multilevel iteration
            r = tts_run(fsb.PC_MLRBM, 5000, 0)
            mi = s.ml_info()
            r.update({"pc": "mlrbm", "ml_levels": mi["levels"], "ml_cells": mi["cells"], "ml_distributed_levels": mi["distributed_levels"], "ml_setup_ms": mi["setup_ms"],
                      "ml_lambda": mi["lambda"]})
            u_ml = s.solution() if args.tts_pc == "both" else None
            tts["multilevel"] = r
        if args.tts_pc in ("jacobi", "both"):
            r = tts_run(fsb.PC_JACOBI, cap, 4096)
            r["pc"] = "jacobi"
            if u_ml is not None and r["converged"]:
                u_j = s.solution()
                r["rel_l2_vs_multilevel"] = float(np.linalg.norm(u_j - u_ml) / np.linalg.norm(u_j))
            tts["jacobi"] = r

    # ---- bounded time-to-solution both arms run (Jacobi-PCG to 1e-8 on a plate the CPU finishes in seconds) ----
    tts_small = None
    if args.tts != "off" and world == 1:
        ms_ = make_workload(fsb.meshgen, TTS_SMALL_NODES, TTS_SMALL_NODES)
        c2 = fsb.FemShell(device=local_rank)
        c2.set_material(NU, EM, THICK)
        c2.set_mesh(ms_["xyz"], ms_["etype"], ms_["eptr"], ms_["enodes"], ms_["bc"])
        c2.set_nodal_loads(ms_["forces"])
        c2.assemble()
        c2.solve(rtol=1e-8, max_its=8, pc=fsb.PC_JACOBI, warm_start=False, allow_not_converged=True)   # schedule + graph capture
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        a_ms = c2.assemble()
        info = c2.solve(rtol=1e-8, max_its=2000000, pc=fsb.PC_JACOBI, warm_start=False, check_every=1024, allow_not_converged=True)
        torch.cuda.synchronize()
        tts_small = {"workload": "meshGen %dx%d nodes Quad-4 (%d DOF), clamped, uniform pressure: assemble + Jacobi-PCG to rtol 1e-8"
                                 % (TTS_SMALL_NODES, TTS_SMALL_NODES, 6 * TTS_SMALL_NODES ** 2),
                     "seconds": time.perf_counter() - t0, "assemble_s": a_ms * 1e-3, "solve_s": info.solve_ms * 1e-3,
                     "iterations": info.iterations, "rel_residual": info.rel_residual, "converged": info.status == 0}
        c2.close()

    # ---- N > 1: world-N oracle parity pre-flight and the north-star plate (BASELINE configs[2]) ----
    dist_parity = c3 = None
    comm_used_mode = s.comm_mode() if world > 1 else None
    if world > 1:
        s.close()
        del s
        if not args.no_parity:
            dist_parity = dist_parity_preflight(fsb, torch, dist, rank, world, local_rank, comm)
        if args.c3 != "off":
            c3 = run_c3(args, fsb, torch, dist, rank, world, local_rank, comm, peak)
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return

    # ---- CPU baseline on the host cores (rank 0): the oracle on a bounded sample, ALL host threads ----
    cpu = None
    if not args.no_cpu:
        from oracle import fso
        threads = host_threads()
        if orc is not None:
            sysm, t_asm, n_cpu_elem, cpu_nx, cpu_ny = orc["sys"], orc["t_asm"], n_elem, nx, ny
        else:                                    # N > 1: one strip (the per-GPU share of the weak-scaled workload)
            mesh1, forces1 = make_workload(fso.meshgen, nx, nx)
            t0 = time.perf_counter()
            sysm = fso.assemble(mesh1, forces1, NU, EM, THICK, threads=threads)
            t_asm, n_cpu_elem, cpu_nx, cpu_ny = time.perf_counter() - t0, mesh1.n_elem, nx, nx
        c_it = args.cpu_iters
        fso.pcg(sysm, rtol=1e-30, max_its=10, threads=threads)   # thread pool, page placement
        t0 = time.perf_counter()
        fso.pcg(sysm, rtol=1e-30, max_its=c_it, threads=threads)
        t_cg = time.perf_counter() - t0
        cpu = {"value": 6 * sysm.n_dofnodes * c_it / t_cg, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "%dx%d-node system%s: full values pass + %d Jacobi-PCG iterations (OpenMP over rows)"
                         % (cpu_nx, cpu_ny, "" if world == 1 else " (one GPU's strip of the weak-scaled workload)", c_it),
               "elements_per_s": n_cpu_elem / t_asm,
               "time_to_solution_bounded": tts_small_reference(fso, threads) if args.tts != "off" else None}

    launches_per_step = 3 + 3 * iters           # rhs, spmv(x0), init, then (spmv+dot, update, direction) per iteration
    if world > 1 and comm_used_mode == fsb.COMM_PEER:
        launches_per_step += 2 + iters          # pack + init finalise, then one halo-push kernel per iteration
    elif world > 1:
        launches_per_step += 2 + 3 * iters      # pack + finalise kernels around the NCCL calls
    traffic, traffic_src = None, None
    if args.traffic is not None:
        traffic, traffic_src = args.traffic, "--traffic flag"
    elif args.nodes == 1000:
        traffic = ncu_traffic(spmv_kernel, None)
        traffic_src = "profiles/ncu_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this kernel on this workload (not re-measured in this run)"
    config = workload_config(nx, ny, world, iters)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": config,
        "details": {"spmv_format": "%d of 36 entries per 6x6 block streamed (%s)" % (fmt["nz_per_block"], "zero-compacted sliced ELL" if fmt["nz_per_block"] < 36 else "parity block-CSR"),
                    "matrix_gb_per_gpu": 1e-9 * fmt["matrix_bytes"], "comm": comm_used, "iterations_counted": its_timed,
                    "spmv_ms_reading_p": spmv_ms_on_p,
                    "peer_wait_us_per_iteration_rank0": None if not waits or not waits["pq_waits"] else
                    {"halo_stamp_in_spmv": waits["halo_wait_us"] / max(1, waits["halo_waits"]), "p_Ap_partials_in_update": waits["pq_wait_us"] / waits["pq_waits"],
                     "r_z_partials_in_direction": waits["rz_wait_us"] / max(1, waits["rz_waits"]),
                     "kernel_wall_us": {"k_spmv_sell": waits["spmv_us"] / waits["pq_waits"], "k_update": waits["update_us"] / waits["pq_waits"],
                                        "k_direction": waits["direction_us"] / max(1, waits["rz_waits"]), "iteration": 1e3 * ms_per_step / iters},
                     "how": "halo: average per waiting warp of k_spmv_sell; partials: clock64 around the spin loops, block 0 (fs_peer.cuh)"},
                    # link-side evidence: NVML's NVLink data counters of rank 0's GPU around the timed region, against the bytes the
                    # algorithm must send per iteration (48 B per send-list node + the 8-byte words of the two mailbox reductions)
                    "nvlink_rank0": ({"available": False, "note": "NVML and nvidia-smi report N/A for the NVLink byte counters on these boxes, and ncu's link "
                                      "metrics need kernel replay, which a kernel that waits for its peer does not survive (tried: profiles/r02p_nvlink_attempt.txt); "
                                      "the device-side wait counters above are the evidence there is"} if world > 1 else None) if not (nvl0 and nvl1) else
                    {"tx_bytes_per_iteration": 1024.0 * (nvl1[0] - nvl0[0]) / max(1, its_timed), "rx_bytes_per_iteration": 1024.0 * (nvl1[1] - nvl0[1]) / max(1, its_timed),
                     "algorithmic_tx_bytes_per_iteration": 48.0 * nx + 16.0 * 3 * (world - 1),   # rank 0: one neighbour, one node row of the strip
                     "source": "NVML NVLINK_THROUGHPUT_DATA_TX/RX, all links of rank 0's GPU, KiB granularity"}},
        "metrics": {"cg_dof_iterations_per_s": value, "elements_assembled_per_s": n_elem / (asm_ms * 1e-3),
                    "assemble_ms": asm_ms, "time_to_solution": tts, "time_to_solution_bounded": tts_small,
                    "time_to_first_solution": first,
                    "setup_s_pattern_colouring_upload": t_setup, "first_values_pass_s_incl_gather_schedule": t_first_assemble,
                    "colors": sz["n_colors"], "assembly_mode": args.asm, "dist_parity": dist_parity,
                    "dist_parity_rel_l2": None if dist_parity is None else dist_parity["rel_l2"], "c3": c3},
        "roofline": {"bound": "hbm", "kernel": spmv_kernel, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "peak_source": peak_src, "bytes_per_launch": actual_b, "csr_equiv_bytes_per_launch": csr_b,
                     "csr_equiv_gbs": csr_b / (spmv_ms * 1e-3) / 1e9, "ms_per_launch": spmv_ms,
                     "share_of_step": spmv_ms * (iters + 1) / ms_per_step,
                     "note": "peak = measured copy bandwidth (half reads, half writes); this kernel is 98 % reads, which HBM serves slightly faster, "
                             "so frac can exceed 1; ncu reports 80 % of the nominal 8 TB/s for it (profiles/r01i_ncu_full.txt)",
                     "traffic": traffic, "traffic_source": traffic_src},
        "assembly_roofline": assembly_roofline,
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(Fh.nbytes if world == 1 else 48 * n_own),
                "d2h_bytes_per_step": int(Sh.nbytes if world == 1 else 48 * n_own),
                "includes": "loads H2D, values re-assembly, %d PCG iterations, displacements D2H%s" % (iters, "" if world == 1 else " (per rank: its own strip of loads in, its own rows of the solution out)")},
        "gpu_launches": launches_per_step * args.steps,
        "clocks": clocks,
    }
    emit_line(json_fd, line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nodes", type=int, default=1000, help="nodes per side of one GPU's strip")
    ap.add_argument("--iters", type=int, default=200, help="PCG iterations per step")
    ap.add_argument("--tts", default="auto", choices=["auto", "on", "off"])
    ap.add_argument("--tts-max-s", type=float, default=90.0)
    ap.add_argument("--tts-pc", default="ml", choices=["ml", "jacobi", "both"], help="preconditioner(s) of the time-to-solution run")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-iters", type=int, default=60)
    ap.add_argument("--ref-nodes", type=int, default=1000, help="reference arm: nodes per side of one strip (same as --nodes)")
    ap.add_argument("--ref-iters", type=int, default=0, help="reference arm: iterations per step; 0 = the config's, bounded by --ref-budget-s")
    ap.add_argument("--ref-budget-s", type=float, default=150.0, help="reference arm: wall-clock bound of the whole steps+warmup loop")
    ap.add_argument("--ref-threads", type=int, default=0, help="reference arm: OpenMP threads; 0 = every core this process may use")
    ap.add_argument("--c3", default="auto", choices=["auto", "on", "off"], help="N > 1: also run BASELINE configs[2] (metrics.c3)")
    ap.add_argument("--c3-nodes", type=int, default=4000)
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the world-N oracle parity pre-flight")
    ap.add_argument("--asm", default="gather", choices=["colored", "gather"])
    ap.add_argument("--comm", default="auto", choices=["auto", "nccl", "peer"], help="multi-GPU exchange inside the CG iteration")
    ap.add_argument("--spmv", default="auto", choices=["auto", "full"], help="full = always stream the parity block-CSR (explicit zeros included)")
    ap.add_argument("--traffic", type=float, default=None, help="dram bytes/launch of k_spmv from an ncu --set full capture")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

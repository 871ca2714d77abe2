#!/usr/bin/env python
"""bench.py -- headline benchmark of the fem-shell hot path on B200 (contract: see DESIGN.md section 6).

    python bench.py --gpus 1 --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --steps K --warmup W    # CPU restatement of the reference (oracle)
    torchrun ... bench.py --gpus N ...                       # one rank per GPU, weak scaling

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


def make_workload(fsb, nodes_x, nodes_y):
    """meshGen: q nx ny 0 0 Lx Ly 1,1,1,1 q0 2 1 z  (clamped = boundary id 1, uniform load)"""
    return fsb.meshgen("q", nodes_x - 1, nodes_y - 1, 0.0, 0.0, PLATE, PLATE * (nodes_y - 1) / (nodes_x - 1),
                       (1, 1, 1, 1), QLOAD, 2, 1)


def spmv_bytes(n_dof, n_blocks, matrix_bytes):
    """bytes one SpMV launch must move: the matrix in the format actually streamed (fs_get_spmv_format:
    values + column ids + row/slice pointers) + x read once + y written once; and the SURVEY.md section 8d
    figure for the same matrix as a scalar CSR with 32-bit column ids (explicit zeros of the 6x6 blocks included)"""
    actual = matrix_bytes + 8 * n_dof + 8 * n_dof
    csr_equiv = 12 * 36 * n_blocks + 20 * n_dof
    return actual, csr_equiv


# ------------------------------------------------------------------------------------------------
def run_reference(args):
    """the reference's CPU implementation of the path, restated (oracle/fs_oracle.c; the reference itself
    needs libMesh+PETSc+MPI, which do not exist in this image) on the box's host cores"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import fso
    import fem_shell_b200 as fsb   # only for the host-side mesh generator (no GPU work on this arm)
    threads = fso.max_threads()
    nodes = args.ref_nodes
    m = make_workload(fsb, nodes, nodes)
    om = fso.Mesh(m["xyz"], m["etype"], m["eptr"], m["enodes"], m["bc"])
    t0 = time.perf_counter()
    sysm = fso.assemble(om, m["forces"], NU, EM, THICK, threads=threads)
    t_asm = time.perf_counter() - t0
    n_dof = 6 * sysm.n_dofnodes
    iters = args.ref_iters
    x = None
    for _ in range(args.warmup):
        x, _, _ = fso.pcg(sysm, rtol=1e-30, max_its=iters, threads=threads)
    t0 = time.perf_counter()
    for k in range(args.steps):
        x, its, _ = fso.pcg(sysm, b=sysm.rhs * (1.0 + np.sin(k / 25.01)), rtol=1e-30, max_its=iters, threads=threads)
    dt = time.perf_counter() - t0
    value = n_dof * iters * args.steps / dt
    sample = "%dx%d-node Quad-4 plate (%d DOF), %d Jacobi-PCG iterations per step, OpenMP CSR" % (nodes, nodes, n_dof, iters)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "meshGen %dx%d nodes DKQ+PLANE Quad-4, clamped, uniform pressure" % (nodes, nodes),
                   "iters_per_step": iters, "pc": "jacobi"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "elements_per_s": om.n_elem / t_asm},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
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
    m = make_workload(fsb, nx, ny)
    n_nodes, n_elem = m["xyz"].shape[0], m["etype"].size
    comm = {"auto": fsb.COMM_AUTO, "nccl": fsb.COMM_NCCL, "peer": fsb.COMM_PEER}[args.comm]
    s = fsb.FemShell(device=local_rank, rank=rank, world=world, nccl_id=nccl_id, comm=comm)
    s.set_material(NU, EM, THICK)
    s.set_assembly_mode(fsb.ASM_GATHER if args.asm == "gather" else fsb.ASM_COLORED)
    s.set_spmv_format(fsb.SPMV_FULL if args.spmv == "full" else fsb.SPMV_AUTO)
    t0 = time.perf_counter()
    s.set_mesh(m["xyz"], m["etype"], m["eptr"], m["enodes"], m["bc"])
    t_setup = time.perf_counter() - t0
    s.set_nodal_loads(m["forces"])
    sz = s.sizes()
    n_dof = 6 * sz["n_dofnodes"]
    n_own = sz["own_end"] - sz["own_begin"]
    stream = torch.cuda.ExternalStream(s.stream, device=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

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

    # ---- assembly: values pass over all elements (pattern + colouring were built once in set_mesh; the row-gather
    # schedule is planned on the host by the first values pass, reported separately) ----
    t0 = time.perf_counter()
    s.assemble()
    t_first_assemble = time.perf_counter() - t0
    for _ in range(max(args.warmup, 3)):
        s.assemble()
    asm_ms = timed(lambda k: s.assemble(), args.steps) / args.steps

    # ---- the timed region: K steps of (rhs + ITERS PCG iterations) ----
    iters = args.iters

    def step(k):
        s.build_rhs(1.0 + np.sin(k / 25.01))
        s.solve(rtol=1e-30, max_its=iters, pc=fsb.PC_JACOBI, warm_start=False, check_every=iters, allow_not_converged=True)

    for k in range(args.warmup):
        step(k)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    total_ms = timed(step, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = total_ms / args.steps
    value = n_dof * iters / (ms_per_step * 1e-3)

    # ---- dominant kernel: SpMV, timed alone on the same stream right after the timed region ----
    spmv_ms = max_over_ranks(s.bench_spmv(50))
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
        "kernel": "k_assemble_gather" if args.asm == "gather" else "k_assemble_colored",
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
            s.solve_host(Fh, Sh, reassemble=True, rtol=1e-30, max_its=iters, pc=fsb.PC_JACOBI, warm_start=False,
                         check_every=iters, allow_not_converged=True)
        else:
            s.solve_host(Fh, None, reassemble=True, rtol=1e-30, max_its=iters, pc=fsb.PC_JACOBI, warm_start=False,
                         check_every=iters, allow_not_converged=True)
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
    tts = None
    if args.tts != "off":
        per_iter_ms = ms_per_step / iters
        cap = int(max(1000, min(args.tts_max_s * 1e3 / per_iter_ms, 5e6)))
        tts = {}

        def tts_run(pc, max_its, check_every):
            s.build_rhs(1.0)
            barrier()
            t0 = time.perf_counter()
            a_ms = s.assemble()
            info = s.solve(rtol=1e-8, max_its=max_its, pc=pc, warm_start=False, check_every=check_every, allow_not_converged=True)
            barrier()
            return {"seconds": time.perf_counter() - t0, "assemble_ms": a_ms, "solve_ms": info.solve_ms, "iterations": info.iterations,
                    "rel_residual": info.rel_residual, "converged": info.status == 0, "rtol": 1e-8, "iteration_cap": max_its}

        if args.tts_pc in ("ml", "both"):
            tts_run(fsb.PC_MLRBM, 4, 0)          # builds the lattice hierarchy (once per mesh) and captures the iteration
            r = tts_run(fsb.PC_MLRBM, 5000, 0)
            mi = s.ml_info()
            r.update({"pc": "mlrbm", "ml_levels": mi["levels"], "ml_cells": mi["cells"], "ml_setup_ms": mi["setup_ms"],
                      "ml_lambda": mi["lambda"]})
            u_ml = s.solution() if args.tts_pc == "both" else None
            tts["multilevel"] = r
        if args.tts_pc in ("jacobi", "both"):
            r = tts_run(fsb.PC_JACOBI, cap, 256)
            r["pc"] = "jacobi"
            if args.tts_pc == "both" and r["converged"]:
                u_j = s.solution()
                r["rel_l2_vs_multilevel"] = float(np.linalg.norm(u_j - u_ml) / np.linalg.norm(u_j))
            tts["jacobi"] = r

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- CPU baseline on the host cores (rank 0, N=1 only): the oracle on a bounded sample ----
    cpu = None
    if world == 1 and not args.no_cpu:
        from oracle import fso
        threads = fso.max_threads()
        om = fso.Mesh(m["xyz"], m["etype"], m["eptr"], m["enodes"], m["bc"])
        t0 = time.perf_counter()
        sysm = fso.assemble(om, m["forces"], NU, EM, THICK, threads=threads)
        t_asm = time.perf_counter() - t0
        c_it = args.cpu_iters
        fso.pcg(sysm, rtol=1e-30, max_its=2, threads=threads)
        t0 = time.perf_counter()
        fso.pcg(sysm, rtol=1e-30, max_its=c_it, threads=threads)
        t_cg = time.perf_counter() - t0
        cpu = {"value": n_dof * c_it / t_cg, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "same %dx%d-node system: full values pass + %d Jacobi-PCG iterations (OpenMP over rows)" % (nx, ny, c_it),
               "elements_per_s": n_elem / t_asm}

    launches_per_step = 3 + 3 * iters           # rhs, spmv(x0), init, then (spmv+dot, update, direction) per iteration
    if world > 1 and s.comm_mode() == fsb.COMM_PEER:
        launches_per_step += 2 + iters          # pack + init finalise, then one halo-push kernel per iteration
    elif world > 1:
        launches_per_step += 2 + 3 * iters      # pack + finalise kernels around the NCCL calls
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "BASELINE configs[1]: meshGen %dx%d nodes DKQ+PLANE Quad-4 (%d elements, %d DOF), clamped, uniform pressure%s"
                   % (nx, ny, n_elem, n_dof, "" if world == 1 else " = one 1000x1000-node strip per GPU"),
                   "iters_per_step": iters, "pc": "jacobi", "dof_order": "first_encounter",
                   "spmv_format": "%d of 36 entries per 6x6 block streamed (%s)" % (fmt["nz_per_block"], "zero-compacted sliced ELL" if fmt["nz_per_block"] < 36 else "parity block-CSR"),
                   "l2_policy": "inputs larger than L2 (matrix %.2f GB per GPU streamed every iteration)" % (1e-9 * fmt["matrix_bytes"]),
                   "parallelism": "node-block strips x%d" % world, "comm": comm_used},
        "metrics": {"cg_dof_iterations_per_s": value, "elements_assembled_per_s": n_elem / (asm_ms * 1e-3),
                    "assemble_ms": asm_ms, "time_to_solution": tts, "setup_s_pattern_colouring_upload": t_setup, "first_values_pass_s_incl_gather_schedule": t_first_assemble,
                    "colors": sz["n_colors"], "assembly_mode": args.asm},
        "roofline": {"bound": "hbm", "kernel": spmv_kernel, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "peak_source": peak_src, "bytes_per_launch": actual_b, "csr_equiv_bytes_per_launch": csr_b,
                     "csr_equiv_gbs": csr_b / (spmv_ms * 1e-3) / 1e9, "ms_per_launch": spmv_ms,
                     "share_of_step": spmv_ms * (iters + 1) / ms_per_step,
                     "note": "peak = measured copy bandwidth (half reads, half writes); this kernel is 98 % reads, which HBM serves slightly faster, "
                             "so frac can exceed 1; ncu reports 80 % of the nominal 8 TB/s for it (profiles/r01i_ncu_full.txt)",
                     "traffic": ncu_traffic(spmv_kernel, args.traffic) if args.nodes == 1000 else args.traffic},
        "assembly_roofline": assembly_roofline,
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(Fh.nbytes if world == 1 else 48 * n_own),
                "d2h_bytes_per_step": int(Sh.nbytes if world == 1 else 48 * n_own),
                "includes": "loads H2D, values re-assembly, %d PCG iterations, displacements D2H%s" % (iters, "" if world == 1 else " (per rank: its own strip of loads in, its own rows of the solution out)")},
        "gpu_launches": launches_per_step * args.steps,
        "clocks": clocks,
    }
    emit_line(json_fd, line)
    if world > 1:
        dist.destroy_process_group()


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
    ap.add_argument("--cpu-iters", type=int, default=30)
    ap.add_argument("--ref-nodes", type=int, default=1000)
    ap.add_argument("--ref-iters", type=int, default=10)
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

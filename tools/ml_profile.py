"""one multilevel-preconditioned solve of the bench workload with the profiler range around a few iterations
(run under: ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv ...)"""
import sys, os, time
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import torch
import fem_shell_b200 as fsb

nodes = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
its = int(sys.argv[2]) if len(sys.argv) > 2 else 3
what = sys.argv[3] if len(sys.argv) > 3 else "iter"      # "iter": a few iterations, "setup": the values set-up
kind = sys.argv[4] if len(sys.argv) > 4 else "q"
m = fsb.meshgen(kind, nodes - 1, nodes - 1, 0.0, 0.0, 10.0, 10.0, (1, 1, 1, 1), 300.0, 2, 1)
s = fsb.FemShell()
s.set_material(0.3, 1.0e7, 0.5)
s.set_mesh(m["xyz"], m["etype"], m["eptr"], m["enodes"], m["bc"])
s.set_nodal_loads(m["forces"])
s.assemble()
s.build_rhs(1.0)
t0 = time.perf_counter()
info = s.solve(rtol=1e-8, max_its=5000, pc=fsb.PC_MLRBM, warm_start=False)
print("first solve: its", info.iterations, "solve_ms", info.solve_ms, "wall", time.perf_counter() - t0, s.ml_info(), flush=True)
rt = torch.cuda.cudart()
if what == "setup":
    s.assemble()
    rt.cudaProfilerStart()
    s.solve(rtol=1e-8, max_its=1, pc=fsb.PC_MLRBM, warm_start=False, allow_not_converged=True)
    torch.cuda.synchronize()
    rt.cudaProfilerStop()
else:
    rt.cudaProfilerStart()
    s.solve(rtol=1e-30, max_its=its, pc=fsb.PC_MLRBM, warm_start=False, check_every=its, allow_not_converged=True)
    torch.cuda.synchronize()
    rt.cudaProfilerStop()

"""time to the first solution of a new mesh, phase by phase (FS_TIMING=1 prints the library's own phases):
   python tools/first_solve_probe.py q 1000   |   t 1000   |   t 4000 (96 M DOF on one GPU: slice pass only)"""
import os, sys, time
os.environ.setdefault("FS_TIMING", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fem_shell_b200 as fsb

kind, nodes = sys.argv[1], int(sys.argv[2])
pc = fsb.PC_MLRBM
t0 = time.perf_counter()
m = fsb.meshgen(kind, nodes - 1, nodes - 1, 0.0, 0.0, 10.0, 10.0, (1, 1, 1, 1), 300.0, 2, 1)
t_gen = time.perf_counter() - t0
s = fsb.FemShell(device=0)
s.set_material(0.3, 1e7, 0.5)
marks = [("meshgen", t_gen)]
def lap(name, fn):
    t0 = time.perf_counter(); r = fn(); marks.append((name, time.perf_counter() - t0)); return r
lap("set_mesh", lambda: s.set_mesh(m["xyz"], m["etype"], m["eptr"], m["enodes"], m["bc"]))
lap("set_nodal_loads", lambda: s.set_nodal_loads(m["forces"]))
lap("first assemble", lambda: s.assemble())
info = lap("first solve (ml set-up, capture, pcg 1e-8)", lambda: s.solve(rtol=1e-8, max_its=5000, pc=pc, warm_start=False))
lap("solution_owned", lambda: s.solution_owned(with_ids=False))
lap("second assemble", lambda: s.assemble())
info2 = lap("second solve", lambda: s.solve(rtol=1e-8, max_its=5000, pc=pc, warm_start=False))
for k, v in marks:
    print("%-45s %9.1f ms" % (k, 1e3 * v))
print("iterations", info.iterations, info2.iterations, "solve_ms", info2.solve_ms, "ml", s.ml_info()["setup_ms"], "fmt", s.spmv_format()["nz_per_block"])

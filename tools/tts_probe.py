"""time-to-solution probe on the bench workload: iterations to rtol for several preconditioner / norm choices"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fem_shell_b200 as fsb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
kind = sys.argv[2] if len(sys.argv) > 2 else "q"
cap = int(sys.argv[3]) if len(sys.argv) > 3 else 600000
m = fsb.meshgen(kind, n - 1, n - 1, 0, 0, 10, 10, (1, 1, 1, 1), 300.0, 2, 1)
s = fsb.FemShell()
s.set_material(0.3, 1e7, 0.5)
s.set_mesh(m["xyz"], m["etype"], m["eptr"], m["enodes"], m["bc"])
s.set_nodal_loads(m["forces"])
s.assemble()
for pc, norm in ((1, 0), (1, 1), (2, 0)):
    t0 = time.perf_counter()
    i = s.solve(rtol=1e-8, max_its=cap, pc=pc, norm_type=norm, warm_start=False, check_every=512, allow_not_converged=True)
    u = s.solution()
    print("n=%d %s pc=%d norm=%d: its=%d rel=%.3e status=%d solve %.2f s  w_center=%.8g" %
          (n, kind, pc, norm, i.iterations, i.rel_residual, i.status, time.perf_counter() - t0, u[(n * n) // 2, 2]), flush=True)

"""times the pieces of the host-buffer plugin call on the bench workload (diagnostic, not a benchmark)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fem_shell_b200 as fsb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
m = fsb.meshgen("q", n - 1, n - 1, 0, 0, 10, 10, (1, 1, 1, 1), 300.0, 2, 1)
s = fsb.FemShell()
s.set_material(0.3, 1e7, 0.5)
t0 = time.perf_counter(); s.set_mesh(m["xyz"], m["etype"], m["eptr"], m["enodes"], m["bc"]); print("set_mesh %.1f ms" % (1e3 * (time.perf_counter() - t0)))
F = torch.from_numpy(m["forces"]).pin_memory().numpy()
S = torch.empty((n * n, 6), dtype=torch.float64).pin_memory().numpy()
for rep in range(3):
    t = [time.perf_counter()]
    s.set_nodal_loads(F); torch.cuda.synchronize(); t.append(time.perf_counter())
    s.assemble(); t.append(time.perf_counter())
    i = s.solve(rtol=1e-30, max_its=200, warm_start=False, check_every=200, allow_not_converged=True); t.append(time.perf_counter())
    s.solution(S); t.append(time.perf_counter())
    print("loads %.2f  assemble %.2f  solve %.2f (device %.2f)  solution %.2f ms" % tuple([1e3 * (t[k + 1] - t[k]) for k in range(3)][:2] + [1e3 * (t[3] - t[2]), i.solve_ms, 1e3 * (t[4] - t[3])]))
t0 = time.perf_counter()
s.solve_host(F, S, reassemble=True, rtol=1e-30, max_its=200, warm_start=False, check_every=200, allow_not_converged=True)
print("solve_host %.2f ms" % (1e3 * (time.perf_counter() - t0)))

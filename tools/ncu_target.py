"""what `ncu --set full --profile-from-start off` should see of the bench workload (c2 by default): one values pass
and one captured batch of 8 Jacobi-PCG iterations, all of them live (no kernel returns early on the done flag)
    ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/TAG python tools/ncu_target.py [nodes] [q|t]"""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import numpy as np
import torch
import fem_shell_b200 as fsb

nodes = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
kind = sys.argv[2] if len(sys.argv) > 2 else "q"
m = fsb.meshgen(kind, nodes - 1, nodes - 1, 0.0, 0.0, 10.0, 10.0, (1, 1, 1, 1), 300.0, 2, 1)
s = fsb.FemShell()
s.set_material(0.3, 1.0e7, 0.5)
s.set_mesh(m["xyz"], m["etype"], m["eptr"], m["enodes"], m["bc"])
s.set_nodal_loads(m["forces"])
for _ in range(3):
    s.assemble()
s.build_rhs(1.0)
kw = dict(rtol=1e-30, pc=fsb.PC_JACOBI, warm_start=False, allow_not_converged=True)
s.solve(max_its=64, check_every=64, **kw)      # capture + warm caches
torch.cuda.synchronize()
rt = torch.cuda.cudart()
rt.cudaProfilerStart()
s.assemble()
info = s.solve(max_its=8, check_every=8, **kw)
torch.cuda.synchronize()
rt.cudaProfilerStop()
print("profiled: 1 values pass (%s), %d iterations" % (s.assembly_path(), info.iterations))

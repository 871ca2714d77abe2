"""SpMV and values pass on the 1000x1000-node Quad-4 plate, axis-aligned vs rotated by 30 degrees
about x (plane frame path) vs the parity format the rotated plate used before"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fem_shell_b200 as fsb

m0 = fsb.meshgen("q", 999, 999, 0.0, 0.0, 10.0, 10.0, (1, 1, 1, 1), 300.0, 2, 1)
c, s_ = np.cos(np.pi / 6), np.sin(np.pi / 6)
R = np.array([[1, 0, 0], [0, c, -s_], [0, s_, c]])
out = {}
for name, rot, full in (("axis_aligned", False, False), ("rotated_30deg_plane_frame", True, False), ("rotated_30deg_parity_format", True, True)):
    m = dict(m0)
    if rot:
        m["xyz"] = np.ascontiguousarray(m0["xyz"] @ R.T)
        m["forces"] = np.ascontiguousarray(np.hstack([m0["forces"][:, :3] @ R.T, m0["forces"][:, 3:] @ R.T]))
    s = fsb.FemShell(device=0)
    s.set_material(0.3, 1e7, 0.5)
    if full:
        s.set_spmv_format(fsb.SPMV_FULL)
    s.set_mesh(m["xyz"], m["etype"], m["eptr"], m["enodes"], m["bc"])
    s.set_nodal_loads(m["forces"])
    for _ in range(3):
        s.assemble()
    asm = min(s.assemble() for _ in range(10))
    spmv = s.bench_spmv(50)
    info = s.solve(rtol=1e-30, max_its=400, pc=fsb.PC_JACOBI, warm_start=False, check_every=400, allow_not_converged=True)
    fmt = s.spmv_format()
    out[name] = {"assemble_ms": asm, "spmv_ms": spmv, "nz_per_block": fmt["nz_per_block"], "matrix_gb": fmt["matrix_bytes"] * 1e-9,
                 "cg_ms_per_iteration": info.solve_ms / info.iterations}
    s.close()
out["spmv_rotated_over_aligned"] = out["rotated_30deg_plane_frame"]["spmv_ms"] / out["axis_aligned"]["spmv_ms"]
print(json.dumps(out))

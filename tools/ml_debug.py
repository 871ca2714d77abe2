"""GPU-box debugging aid: lattice levels of FS_PC_MLRBM against the explicit-Galerkin mirror, level by level."""
import sys, os
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import fem_shell_b200 as fsb
from oracle import fso
from ml_mirror import Mirror, element_extent, stencil_to_csr, diag_blocks, pinv_blocks

m = fsb.meshgen("q", 40, 33, 0, 0, 10, 8.25, (1, 1, 1, 1), 300.0, 2, 1)
om = fso.Mesh(np.asarray(m["xyz"], float), m["etype"], m["eptr"], m["enodes"], m["bc"])
ref = fso.assemble(om, m["forces"], 0.3, 1e7, 0.5)
s = fsb.FemShell(); s.set_material(0.3, 1e7, 0.5); s.set_mesh(m["xyz"], m["etype"], m["eptr"], m["enodes"], m["bc"])
s.set_nodal_loads(m["forces"]); s.assemble(); s.set_ml_options(dense_points=24, gamma=1)
r = np.random.default_rng(7).standard_normal(6 * ref.n_dofnodes)
z = s.apply_mlrbm(r)
info = s.ml_info(); print(info)
xyz = np.zeros((ref.n_dofnodes, 3)); xyz[ref.dofnode] = m["xyz"]; mask = np.zeros(ref.n_dofnodes, np.uint8); mask[ref.dofnode] = ref.mask
h = element_extent(np.asarray(m["xyz"], float), m["eptr"], m["enodes"])
lam = list(info["lambda"])
lam_use = [lam[0]] + [2.2] * (len(lam) - 1)
M = Mirror(ref.scipy().tocsr(), xyz, mask, h, info["cells"], lam_use, gamma=1)
active = [1, 1, 0]
for l in range(info["levels"]):
    Ag = stencil_to_csr(s.ml_level(l, 0), info["cells"][l], active)
    lev = M.levels[l + 1] if l + 1 < len(M.levels) else None
    # mirror matrix of lattice l = A of mirror level l+1 (or the dense one)
    if l + 1 < len(M.levels) and "A" in M.levels[l + 1]:
        Am = M.levels[l + 1]["A"]
        d = (Ag - Am)
        print("level", l, "|A_gpu - A_mirror|/|A_mirror| =", abs(d).max() / abs(Am).max(), "sym err gpu", abs(Ag - Ag.T).max() / abs(Ag).max())
        if l == 0:
            n = Ag.shape[0] // 6
            Ab = Ag.toarray(); Bm = Am.toarray()
            # where do they differ: print block (cell 7) rows
            c = (info["cells"][0][0] * 3 + 3)
            print("gpu diag block cell", c); print(np.array2string(Ab[6*c:6*c+6, 6*c:6*c+6], precision=3))
            print("mirror diag block"); print(np.array2string(Bm[6*c:6*c+6, 6*c:6*c+6], precision=3))
            print("gpu right block"); print(np.array2string(Ab[6*c:6*c+6, 6*c+6:6*c+12], precision=3))
            print("mirror right block"); print(np.array2string(Bm[6*c:6*c+6, 6*c+6:6*c+12], precision=3))
        Dg = s.ml_level(l, 1).reshape(-1, 6, 6)
        Dm = pinv_blocks(diag_blocks(Am, Am.shape[0] // 6))
        print("   dinv diff", abs(Dg - Dm).max() / abs(Dm).max())
        per = abs(Dg - Dm).reshape(Dg.shape[0], -1).max(1)
        bad = np.argsort(-per)[:4]
        print("   worst cells", bad, per[bad], "of", Dg.shape[0], "cells;  #cells with diff > 1e-9*max:", int((per > 1e-9 * abs(Dm).max()).sum()))
        wc = bad[0]
        np.set_printoptions(linewidth=200)
        print("   A block of worst cell (gpu)"); print(np.array2string(Ag[6*wc:6*wc+6, 6*wc:6*wc+6].toarray(), precision=4))
        print("   gpu dinv"); print(np.array2string(Dg[wc], precision=4))
        print("   mirror dinv"); print(np.array2string(Dm[wc], precision=4))
        import scipy.sparse as sp
        from ml_mirror import block_diag
        for nm, DD in (("gpu", Dg), ("mirror", Dm)):
            Di = block_diag(DD)
            v = np.random.default_rng(0).standard_normal(Ag.shape[0])
            for _ in range(100):
                w = Di @ (Ag @ v); lamv = np.linalg.norm(w) / np.linalg.norm(v); v = w / np.linalg.norm(w)
            print("   lambda_max(D^+ A_gpu) with", nm, "dinv:", lamv)
    else:
        print("level", l, "dense")
zr = M(r)
print("cycle diff (mirror uses lambda 2.2 on lattices)", np.linalg.norm(z - zr) / np.linalg.norm(zr))

"""Regenerates tests/golden/ref_meshes.npz and tests/golden/meshgen_ref.npz from the
read-only reference tree.  Run in the build container only (it reads /root/reference and
runs oracle/_ref/meshgen, the reference's own generator compiled from its source):

    python tests/golden/make_fixtures.py

ref_meshes.npz   the reference's shipped example inputs (src/fem-shell/example-meshes/*.xda,
                 *_f and preCICE/example-meshes/bending_tower_tri_test.xda) as arrays, so the
                 GPU box -- where /root/reference does not exist -- can run the thesis cases.
meshgen_ref.npz  outputs of the reference meshGen binary for a handful of argument sets, used
                 to pin the numpy / C++ restatements of the generator (incl. %g rounding).
The thesis' published displacements live in thesis_goldens.json (hand-transcribed from
doc/validation.tex, line numbers inside).
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import fso  # noqa: E402

REF = "/root/reference/src/fem-shell"
OUT = os.path.dirname(os.path.abspath(__file__))

MESHGEN_CASES = {
    # name: (kind, nx, ny, min_x, min_y, max_x, max_y, bcids(t,b,l,r), factor, loading, ul_lr, dead)
    "q7x5_uni": ("q", 7, 5, 0.0, 0.0, 10.0, 3.0, (1, 0, 20, 21), 117.0, 2, 1, "z"),
    "t9x9_conc_ul": ("t", 9, 9, 0.0, 0.0, 10.0, 10.0, (0, 0, 0, 0), 30000.0, 1, 1, "z"),
    "t6x4_uni_ur": ("t", 6, 4, -1.0, 0.5, 2.0, 3.0, (1, -1, 2, 0), 2.5, 2, 0, "y"),
    "q3x3_none": ("q", 3, 3, 0.0, 0.0, 1.0, 1.0, (-1, 1, -1, -1), 1.0, 0, 1, "x"),
    "q999_strip": ("q", 999, 2, 0.0, 0.0, 10.0, 10.0, (1, 1, 1, 1), 300.0, 2, 1, "z"),
}


def main():
    d = {}
    ex = os.path.join(REF, "example-meshes")
    for name in sorted(os.listdir(ex)):
        if not name.endswith(".xda"):
            continue
        base = name[:-4]
        m = fso.read_xda(os.path.join(ex, name))
        F = fso.read_forces(os.path.join(ex, base + "_f"), m.n_nodes)
        for k, v in dict(xyz=m.xyz, etype=m.etype, eptr=m.eptr, enodes=m.enodes, bc=m.bc, forces=F).items():
            d[base + "/" + k] = v
    m = fso.read_xda(os.path.join(REF, "preCICE/example-meshes/bending_tower_tri_test.xda"))
    for k, v in dict(xyz=m.xyz, etype=m.etype, eptr=m.eptr, enodes=m.enodes, bc=m.bc).items():
        d["bending_tower_tri_test/" + k] = v
    np.savez_compressed(os.path.join(OUT, "ref_meshes.npz"), **d)

    g = {}
    with tempfile.TemporaryDirectory() as tmp:
        for name, a in MESHGEN_CASES.items():
            m, F = fso.run_ref_meshgen(os.path.join(tmp, name), *a)
            for k, v in dict(xyz=m.xyz, etype=m.etype, eptr=m.eptr, enodes=m.enodes, bc=m.bc, forces=F).items():
                g[name + "/" + k] = v
    np.savez_compressed(os.path.join(OUT, "meshgen_ref.npz"), **g)
    print("wrote", len(d), "+", len(g), "arrays")


if __name__ == "__main__":
    main()

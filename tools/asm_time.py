"""values-pass timing of the Tri-3 plate (1000x1000 nodes) and a mixed mesh, one GPU"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fem_shell_b200 as fsb

for kind, n in (("t", 1000), ("q", 1000)):
    m = fsb.meshgen(kind, n - 1, n - 1, 0.0, 0.0, 10.0, 10.0, (1, 1, 1, 1), 300.0, 2, 1)
    s = fsb.FemShell(device=0)
    s.set_material(0.3, 1e7, 0.5)
    s.set_mesh(m["xyz"], m["etype"], m["eptr"], m["enodes"], m["bc"])
    for _ in range(3):
        s.assemble()
    ms = min(s.assemble() for _ in range(10))
    print("%s %dx%d nodes: %d elements, values pass %.3f ms = %.1f M elements/s" % (kind, n, n, m["etype"].size, ms, m["etype"].size / ms / 1e3))
    s.close()

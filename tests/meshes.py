"""Deterministic synthetic meshes for the parity tests (no RNG beyond fixed seeds).

folded_cantilever: BASELINE config 5 -- a cantilever strip folded into an L section, alternating
rows of Quad-4 and Tri-3 cells, rigidly rotated by a fixed non-axis-aligned rotation, clamped
(boundary id 1) at one end, point loads at the tip as in the reference's Test E load file.
`skew` > 0 shears/tapers the cells so that triangles become scalene and quads trapezoidal, which
switches on the reference's arithmetic quirks."""
import numpy as np

TRI3, QUAD4 = 3, 5


def rotation(ax=0.3, ay=-0.5, az=0.7):
    cx, sx, cy, sy, cz, sz = np.cos(ax), np.sin(ax), np.cos(ay), np.sin(ay), np.cos(az), np.sin(az)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def folded_cantilever(nx=12, ny=6, Lx=40.0, Ly=10.0, skew=0.0, rotate=True, tri_rows="alternate"):
    W = nx + 1
    s = np.linspace(0.0, Lx, nx + 1)
    t = np.linspace(0.0, Ly, ny + 1)
    S, T = np.meshgrid(s, t)
    if skew:
        S = S + skew * (Lx / nx) * np.sin(1.3 * T + 0.4) * (S > 0)
        T = T + skew * (Ly / ny) * 0.5 * np.sin(0.9 * S) * ((T > 0) & (T < Ly))
    fold = Ly / 2.0
    y = np.where(T <= fold, T, fold)
    z = np.where(T <= fold, 0.0, T - fold)
    xyz = np.stack([S.ravel(), y.ravel(), z.ravel()], 1)
    if rotate:
        xyz = xyz @ rotation().T + np.array([1.5, -2.0, 0.75])
    etype, enodes, eptr, bc = [], [], [0], []
    for j in range(ny):
        use_tri = (tri_rows == "all") or (tri_rows == "alternate" and j % 2 == 1)
        for i in range(nx):
            A = i + j * W
            B, Cn, D = A + 1, A + W + 1, A + W
            if use_tri:
                for tri in ((A, B, D), (B, Cn, D)):
                    if i == 0 and tri[0] == A:
                        bc.append((len(etype), 2, 1))   # side D->A lies on the clamped end
                    etype.append(TRI3)
                    enodes.extend(tri)
                    eptr.append(len(enodes))
            else:
                if i == 0:
                    bc.append((len(etype), 3, 1))
                etype.append(QUAD4)
                enodes.extend((A, B, Cn, D))
                eptr.append(len(enodes))
    n_nodes = xyz.shape[0]
    F = np.zeros((n_nodes, 6))
    tip_lo, tip_hi = nx, nx + ny * W          # the two free corners of the tip
    F[tip_lo, :3] = (0.0, 1.6, 0.0)
    F[tip_hi, :3] = (0.0, -1.6, 0.4)
    if rotate:
        R = rotation()
        F[:, :3] = F[:, :3] @ R.T
    return dict(xyz=xyz, etype=np.array(etype, np.int32), eptr=np.array(eptr, np.int64),
                enodes=np.array(enodes, np.int32), bc=np.array(bc, np.int32).reshape(-1, 3), forces=F)


def random_elements(n, seed=1234):
    """general (scalene / non-parallelogram, arbitrarily oriented) single elements"""
    rng = np.random.default_rng(seed)
    out = []
    for k in range(n):
        R = rotation(*rng.uniform(-3, 3, 3))
        scale = 10.0 ** rng.uniform(-2, 1)
        if k % 2 == 0:
            P = np.array([[0, 0, 0], [1.0, 0, 0], [0.3, 0.9, 0]]) + rng.uniform(-0.2, 0.2, (3, 3)) * [1, 1, 0]
            et = TRI3
        else:
            P = np.array([[0, 0, 0], [1.2, 0, 0], [1.3, 1.0, 0], [0.1, 0.8, 0]]) + rng.uniform(-0.15, 0.15, (4, 3)) * [1, 1, 0]
            et = QUAD4
        out.append((et, (P * scale) @ R.T + rng.uniform(-5, 5, 3)))
    return out


def elements_as_mesh(elems):
    xyz, etype, enodes, eptr = [], [], [], [0]
    for et, P in elems:
        base = len(xyz)
        xyz.extend(P.tolist())
        etype.append(et)
        enodes.extend(range(base, base + len(P)))
        eptr.append(len(enodes))
    return dict(xyz=np.array(xyz), etype=np.array(etype, np.int32), eptr=np.array(eptr, np.int64),
                enodes=np.array(enodes, np.int32), bc=np.zeros((0, 3), np.int32))


def umbrella(n_spokes=14, n_rings=4, cone=0.35, mixed=True):
    """Unstructured shell around a hub: `n_spokes` triangles meet at node 0 (all with the hub as their FIRST local
    node, so every pair of them clashes in the gather schedule's step 0), followed by rings of quads (or, with
    `mixed`, alternating rings of quads and triangle pairs) on a shallow cone.  Outer ring clamped (id 1), a point
    load and a moment at the hub.  High valence + identical local indices = worst case for the emit phases."""
    xyz = [(0.0, 0.0, 0.0)]
    ring = lambda r: [1 + (r - 1) * n_spokes + k for k in range(n_spokes)]
    for r in range(1, n_rings + 1):
        for k in range(n_spokes):
            a = 2 * np.pi * (k + 0.13 * r) / n_spokes
            rad = r * (1.0 + 0.07 * np.sin(3 * a))
            xyz.append((rad * np.cos(a), rad * np.sin(a), cone * rad))
    etype, enodes, eptr, bc = [], [], [0], []
    r1 = ring(1)
    for k in range(n_spokes):
        etype.append(TRI3)
        enodes.extend((0, r1[k], r1[(k + 1) % n_spokes]))
        eptr.append(len(enodes))
    for r in range(1, n_rings):
        a, b = ring(r), ring(r + 1)
        for k in range(n_spokes):
            k1 = (k + 1) % n_spokes
            if mixed and r % 2 == 0:
                for tri in ((a[k], b[k], a[k1]), (a[k1], b[k], b[k1])):
                    if r == n_rings - 1 and tri[2] == b[k1]:
                        bc.append((len(etype), 1, 1))       # side b[k] -> b[k1] lies on the outer ring
                    etype.append(TRI3)
                    enodes.extend(tri)
                    eptr.append(len(enodes))
            else:
                if r == n_rings - 1:
                    bc.append((len(etype), 1, 1))           # side b[k] -> b[k1]
                etype.append(QUAD4)
                enodes.extend((a[k], b[k], b[k1], a[k1]))
                eptr.append(len(enodes))
    xyz = np.array(xyz) @ rotation(0.2, 0.4, -0.3).T
    F = np.zeros((xyz.shape[0], 6))
    F[0] = (0.3, -0.2, -5.0, 0.4, 0.1, 0.0)
    return dict(xyz=xyz, etype=np.array(etype, np.int32), eptr=np.array(eptr, np.int64),
                enodes=np.array(enodes, np.int32), bc=np.array(bc, np.int32).reshape(-1, 3), forces=F)


def delaunay_patch(n_points=400, seed=7, quad_fraction=0.3):
    """Irregular planar mesh: Delaunay triangulation of seeded random points, a fraction of edge-adjacent triangle pairs
    merged into quads (convex ones only).  Valence 3..10, arbitrary local node indices -- an unstructured input for the
    host-side schedules.  No boundary records, no loads."""
    from scipy.spatial import Delaunay
    rng = np.random.default_rng(seed)
    pts = rng.uniform(0.0, 10.0, (n_points, 2))
    tri = Delaunay(pts).simplices.astype(np.int32)
    # counter-clockwise triangles
    a, b, c = pts[tri[:, 0]], pts[tri[:, 1]], pts[tri[:, 2]]
    cw = (b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (b[:, 1] - a[:, 1]) * (c[:, 0] - a[:, 0]) < 0
    tri[cw] = tri[cw][:, [0, 2, 1]]
    edge_owner, used, quads = {}, np.zeros(len(tri), bool), []
    for t, (i, j, k) in enumerate(tri):
        for e in ((i, j), (j, k), (k, i)):
            edge_owner[e] = t
    for t, (i, j, k) in enumerate(tri):
        if used[t] or rng.uniform() > quad_fraction:
            continue
        for (p, q, r) in ((i, j, k), (j, k, i), (k, i, j)):      # edge p->q, opposite node r
            o = edge_owner.get((q, p))
            if o is None or used[o] or o == t:
                continue
            s = [v for v in tri[o] if v != p and v != q][0]
            quad = [r, p, s, q]                                  # counter-clockwise around the merged pair
            P = pts[quad]
            cross = [(P[(m + 1) % 4, 0] - P[m, 0]) * (P[(m + 2) % 4, 1] - P[(m + 1) % 4, 1])
                     - (P[(m + 1) % 4, 1] - P[m, 1]) * (P[(m + 2) % 4, 0] - P[(m + 1) % 4, 0]) for m in range(4)]
            if min(cross) <= 1e-9:
                continue
            used[t] = used[o] = True
            quads.append(quad)
            break
    etype, enodes, eptr = [], [], [0]
    for t in range(len(tri)):
        if not used[t]:
            etype.append(TRI3); enodes.extend(int(v) for v in tri[t]); eptr.append(len(enodes))
    for qd in quads:
        etype.append(QUAD4); enodes.extend(int(v) for v in qd); eptr.append(len(enodes))
    xyz = np.column_stack([pts, np.zeros(n_points)])
    return dict(xyz=xyz, etype=np.array(etype, np.int32), eptr=np.array(eptr, np.int64), enodes=np.array(enodes, np.int32),
                bc=np.zeros((0, 3), np.int32), forces=np.zeros((n_points, 6)))

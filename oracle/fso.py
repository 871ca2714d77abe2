"""ctypes front-end of the CPU oracle (oracle/fs_oracle.c) plus numpy
restatements of the reference's file formats and mesh generator.

TEST INFRASTRUCTURE ONLY.  May be imported by tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs; never by the product
package fem_shell_b200.

Reference anchors:
  XDA mesh files            src/meshgen/main_all.cpp:233-339 (writer), fs.cpp:37 (reader = libMesh)
  <basename>_f load files   fs.cpp:44-67 (reader), src/meshgen/main_all.cpp:343-387 (writer)
  meshGen                   src/meshgen/main_all.cpp:15-389
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libfs_oracle.so")
REF_MESHGEN = os.path.join(HERE, "_ref", "meshgen")

TRI3, QUAD4 = 3, 5
QUIRKS_REFERENCE = 3
DOF_FIRST_ENCOUNTER, DOF_NODE_ID = 0, 1
PC_NONE, PC_JACOBI, PC_BJACOBI6 = 0, 1, 2

_lib = None


def build():
    subprocess.check_call(["make", "-C", HERE, "--no-print-directory"], stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = C.CDLL(LIB_PATH)
        _lib.fso_dof_order.restype = C.c_int64
        _lib.fso_node_pattern.restype = C.c_int64
        _lib.fso_pcg.restype = C.c_int64
        _lib.fso_element_stiffness.restype = C.c_int
        _lib.fso_max_threads.restype = C.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@dataclass
class Mesh:
    xyz: np.ndarray      # (n_nodes, 3) float64
    etype: np.ndarray    # (n_elem,) int32, 3 = TRI3, 5 = QUAD4 (XDA type ids)
    eptr: np.ndarray     # (n_elem+1,) int64
    enodes: np.ndarray   # (eptr[-1],) int32
    bc: np.ndarray       # (n_bc, 3) int32 rows (element, side, boundary id)

    @property
    def n_nodes(self):
        return self.xyz.shape[0]

    @property
    def n_elem(self):
        return self.etype.shape[0]


# ----------------------------------------------------------------------------
# file formats
# ----------------------------------------------------------------------------
def read_xda(path) -> Mesh:
    """libMesh-0.7.0+ ASCII XDA as written by meshGen (main_all.cpp:233-339)."""
    with open(path) as f:
        lines = [ln.split("#")[0].strip() for ln in f]
    assert lines[0].startswith("libMesh"), "not an XDA file"
    n_elem = int(lines[1])
    n_nodes = int(lines[2])
    pos = 8  # 5 spec lines + level-0 header
    etype = np.empty(n_elem, np.int32)
    eptr = np.zeros(n_elem + 1, np.int64)
    en = []
    for e in range(n_elem):
        t = lines[pos + e].split()
        etype[e] = int(t[0])
        ids = [int(v) for v in t[1:]]
        en.extend(ids)
        eptr[e + 1] = eptr[e] + len(ids)
    pos += n_elem
    xyz = np.array([[float(v) for v in lines[pos + i].split()] for i in range(n_nodes)], np.float64)
    pos += n_nodes
    n_bc = int(lines[pos])
    bc = np.array([[int(v) for v in lines[pos + 1 + i].split()] for i in range(n_bc)], np.int32).reshape(n_bc, 3)
    return Mesh(xyz.reshape(n_nodes, 3), etype, eptr, np.array(en, np.int32), bc)


def read_forces(path, n_nodes=None) -> np.ndarray:
    """fs.cpp:52-66: header n, global factor, then n rows of six numbers.  A
    short file (meshGen writes n-1 rows, main_all.cpp:352,377) leaves the
    remaining rows zero because the failed stream extraction keeps the
    zero-initialised DenseVector."""
    toks = open(path).read().split()
    n = int(toks[0])
    factor = float(toks[1])
    vals = np.array([float(t) for t in toks[2:2 + 6 * n]], np.float64)
    out = np.zeros(6 * n, np.float64)
    out[: vals.size] = vals
    out = out.reshape(n, 6) * factor
    if n_nodes is not None and n < n_nodes:
        out = np.vstack([out, np.zeros((n_nodes - n, 6))])
    return out


def _g6(v):
    """default std::ostream formatting of a double (6 significant digits) and back"""
    return float("%g" % v)


def meshgen(kind, nx, ny, min_x, min_y, max_x, max_y, bcids, factor, loading, ul_lr, dead_axis="z"):
    """numpy restatement of meshGen (main_all.cpp:133-387) INCLUDING the trip
    through 6-significant-digit text that its output files impose.
    bcids = (top, bottom, left, right), -1 for none.  Returns (Mesh, forces)."""
    kind = kind.lower()
    n_nodes = (nx + 1) * (ny + 1)
    fracx = (max_x - min_x) / float(nx)
    fracy = (max_y - min_y) / float(ny)
    xs = np.array([_g6(min_x + x * fracx) for x in range(nx + 1)])
    ys = np.array([_g6(min_y + y * fracy) for y in range(ny + 1)])
    xyz = np.zeros((n_nodes, 3))
    pa = 1 if dead_axis == "x" else 0           # primary axis (main_all.cpp:153-156)
    sa = 1 if dead_axis == "z" else 2           # secondary axis (main_all.cpp:147-150)
    X, Y = np.meshgrid(xs, ys)                  # node id = x + y*(nx+1)
    xyz[:, pa] = X.ravel()
    xyz[:, sa] = Y.ravel()
    xi, yi = np.meshgrid(np.arange(nx), np.arange(ny))
    nid = (xi + yi * (nx + 1)).ravel()
    if kind == "q":
        en = np.stack([nid, nid + 1, nid + nx + 2, nid + nx + 1], 1)
        etype = np.full(nx * ny, QUAD4, np.int32)
        eptr = np.arange(nx * ny + 1, dtype=np.int64) * 4
    else:
        if ul_lr:
            t1 = np.stack([nid, nid + 1, nid + nx + 1], 1)
            t2 = np.stack([nid + 1, nid + nx + 2, nid + nx + 1], 1)
        else:
            t1 = np.stack([nid, nid + nx + 2, nid + 1], 1)
            t2 = np.stack([nid + nx + 2, nid, nid + nx + 1], 1)
        en = np.stack([t1, t2], 1).reshape(-1, 3)
        etype = np.full(2 * nx * ny, TRI3, np.int32)
        eptr = np.arange(2 * nx * ny + 1, dtype=np.int64) * 3
    t_id, b_id, l_id, r_id = bcids
    bc = []
    for i in range(nx):                         # main_all.cpp:284-310
        if kind == "t":
            if ul_lr:
                if b_id >= 0: bc.append((2 * i, 0, b_id))
                if t_id >= 0: bc.append((2 * nx * ny - 2 * i - 1, 1, t_id))
            else:
                if b_id >= 0: bc.append((2 * i, 2, b_id))
                if t_id >= 0: bc.append((2 * nx * ny - 2 * i - 1, 2, t_id))
        else:
            if b_id >= 0: bc.append((i, 0, b_id))
            if t_id >= 0: bc.append((nx * ny - 1 - i, 2, t_id))
    for i in range(ny):                         # main_all.cpp:312-338
        if kind == "t":
            if ul_lr:
                if l_id >= 0: bc.append((2 * nx * i, 2, l_id))
                if r_id >= 0: bc.append((2 * nx * (i + 1) - 1, 0, r_id))
            else:
                if l_id >= 0: bc.append((2 * nx * i + 1, 1, l_id))
                if r_id >= 0: bc.append((2 * nx * (i + 1) - 2, 1, r_id))
        else:
            if l_id >= 0: bc.append((nx * i, 3, l_id))
            if r_id >= 0: bc.append((nx * (i + 1) - 1, 1, r_id))
    mesh = Mesh(xyz, etype, eptr, en.astype(np.int32).ravel(), np.array(bc, np.int32).reshape(-1, 3))
    forces = np.zeros((n_nodes, 6))
    comp = {"x": 0, "y": 1, "z": 2}[dead_axis]
    if loading == 1:                            # main_all.cpp:349-366
        if n_nodes // 2 < n_nodes - 1:
            forces[n_nodes // 2, comp] = 1.0
        forces *= _g6(factor)
    elif loading == 2:                          # main_all.cpp:367-386 (last node gets no row)
        forces[: n_nodes - 1, comp] = 1.0
        forces *= _g6(factor * fracx * fracy)
    return mesh, forces


def write_xda(path, mesh: Mesh):
    with open(path, "w") as f:
        f.write("libMesh-0.7.0+\n%d      # number of elements\n%d      # number of nodes\n" % (mesh.n_elem, mesh.n_nodes))
        f.write(".        # boundary condition specification file\nn/a      # subdomain id specification file\n")
        f.write("n/a      # processor id specification file\nn/a      # p-level specification file\n")
        f.write("%d      # n_elem at level 0, [ type (n0 ... nN-1) ]\n" % mesh.n_elem)
        for e in range(mesh.n_elem):
            ids = mesh.enodes[mesh.eptr[e]:mesh.eptr[e + 1]]
            f.write("%d %s\n" % (mesh.etype[e], " ".join(str(int(i)) for i in ids)))
        for p in mesh.xyz:
            f.write("%g %g %g\n" % tuple(p))
        f.write("%d        # number of boundary conditions\n" % len(mesh.bc))
        for r in mesh.bc:
            f.write("%d %d %d\n" % tuple(r))


def run_ref_meshgen(outbase, kind, nx, ny, min_x, min_y, max_x, max_y, bcids, factor, loading, ul_lr, dead_axis="z"):
    """run the reference's own generator (oracle/_ref/meshgen, built from /root/reference)"""
    args = [REF_MESHGEN, kind, str(nx), str(ny), repr(min_x), repr(min_y), repr(max_x), repr(max_y),
            ",".join(str(b) for b in bcids), repr(factor), str(loading), "1" if ul_lr else "0", dead_axis, outbase]
    subprocess.check_call(args)
    mesh = read_xda(outbase + ".xda")
    forces = read_forces(outbase + "_f", mesh.n_nodes) if loading > 0 else np.zeros((mesh.n_nodes, 6))
    return mesh, forces


# ----------------------------------------------------------------------------
# oracle calls
# ----------------------------------------------------------------------------
def element_stiffness(etype, xyz, nu, em, t, quirks=QUIRKS_REFERENCE, layout=0):
    nen = 3 if etype == TRI3 else 4
    K = np.zeros((6 * nen, 6 * nen))
    X = np.ascontiguousarray(xyz, np.float64)
    lib().fso_element_stiffness(C.c_int(etype), _p(X), C.c_double(nu), C.c_double(em), C.c_double(t),
                                C.c_int(quirks), C.c_int(layout), _p(K))
    return K


def dof_order(mesh: Mesh, mode=DOF_FIRST_ENCOUNTER):
    dn = np.empty(mesh.n_nodes, np.int32)
    n = lib().fso_dof_order(C.c_int64(mesh.n_nodes), C.c_int64(mesh.n_elem), _p(mesh.eptr), _p(mesh.enodes),
                            C.c_int(mode), _p(dn))
    return dn, int(n)


def constraint_mask(mesh: Mesh):
    m = np.zeros(mesh.n_nodes, np.uint8)
    bc = np.ascontiguousarray(mesh.bc, np.int32)
    lib().fso_constraint_mask(C.c_int64(mesh.n_nodes), _p(mesh.eptr), _p(mesh.enodes), C.c_int64(len(bc)), _p(bc), _p(m))
    return m


def node_pattern(mesh: Mesh, dofnode, n_dofnodes):
    nptr = np.zeros(n_dofnodes + 1, np.int64)
    nnzb = lib().fso_node_pattern(C.c_int64(n_dofnodes), C.c_int64(mesh.n_elem), _p(mesh.eptr), _p(mesh.enodes),
                                  _p(dofnode), _p(nptr), None)
    nadj = np.zeros(max(int(nnzb), 1), np.int32)
    lib().fso_node_pattern(C.c_int64(n_dofnodes), C.c_int64(mesh.n_elem), _p(mesh.eptr), _p(mesh.enodes),
                           _p(dofnode), _p(nptr), _p(nadj))
    return nptr, nadj[: int(nnzb)]


def expand_csr(nptr, nadj, with_cols=True):
    nn = nptr.size - 1
    rowptr = np.zeros(6 * nn + 1, np.int64)
    colidx = np.zeros(36 * int(nptr[-1]), np.int32) if with_cols else None
    lib().fso_expand_csr(C.c_int64(nn), _p(nptr), _p(nadj), _p(rowptr), _p(colidx) if with_cols else None)
    return rowptr, colidx


@dataclass
class System:
    dofnode: np.ndarray
    n_dofnodes: int
    mask: np.ndarray
    nptr: np.ndarray
    nadj: np.ndarray
    vals: np.ndarray
    rhs: np.ndarray

    def csr(self):
        rowptr, colidx = expand_csr(self.nptr, self.nadj)
        return rowptr, colidx, self.vals

    def scipy(self):
        import scipy.sparse as sp
        rowptr, colidx, vals = self.csr()
        n = 6 * self.n_dofnodes
        return sp.csr_matrix((vals, colidx, rowptr), shape=(n, n))


def assemble(mesh: Mesh, forces, nu, em, t, quirks=QUIRKS_REFERENCE, dof_mode=DOF_FIRST_ENCOUNTER, threads=1) -> System:
    dn, nn = dof_order(mesh, dof_mode)
    mask = constraint_mask(mesh)
    nptr, nadj = node_pattern(mesh, dn, nn)
    vals = np.zeros(36 * int(nptr[-1]))
    rhs = np.zeros(6 * nn)
    F = None if forces is None else np.ascontiguousarray(forces, np.float64)
    xyz = np.ascontiguousarray(mesh.xyz, np.float64)
    lib().fso_assemble(C.c_int64(mesh.n_nodes), _p(xyz), C.c_int64(mesh.n_elem), _p(mesh.etype), _p(mesh.eptr),
                       _p(mesh.enodes), _p(dn), _p(mask), _p(F) if F is not None else None,
                       C.c_double(nu), C.c_double(em), C.c_double(t), C.c_int(quirks), C.c_int64(nn),
                       _p(nptr), _p(nadj), _p(vals), _p(rhs), C.c_int(threads))
    return System(dn, nn, mask, nptr, nadj, vals, rhs)


def pcg(sysm: System, b=None, x0=None, pc=PC_JACOBI, norm_type=0, rtol=1e-8, max_its=1000000, threads=1):
    b = sysm.rhs if b is None else np.ascontiguousarray(b, np.float64)
    x = np.zeros(6 * sysm.n_dofnodes) if x0 is None else np.array(x0, np.float64)
    rel = C.c_double(0.0)
    its = lib().fso_pcg(C.c_int64(sysm.n_dofnodes), _p(sysm.nptr), _p(sysm.nadj), _p(sysm.vals), _p(b), _p(x),
                        C.c_int(pc), C.c_int(norm_type), C.c_double(rtol), C.c_int64(max_its), C.c_int(threads),
                        C.byref(rel))
    return x, int(its), rel.value


def spmv(sysm: System, x, threads=1):
    y = np.zeros_like(x)
    lib().fso_spmv(C.c_int64(sysm.n_dofnodes), _p(sysm.nptr), _p(sysm.nadj), _p(sysm.vals),
                   _p(np.ascontiguousarray(x, np.float64)), _p(y), C.c_int(threads))
    return y


def gather_solution(mesh: Mesh, sysm: System, x):
    sols = np.zeros(6 * mesh.n_nodes)
    lib().fso_gather_solution(C.c_int64(mesh.n_nodes), _p(sysm.dofnode), _p(np.ascontiguousarray(x)), _p(sols))
    return sols.reshape(-1, 6)


def direct_solve(mesh: Mesh, sysm: System):
    """sparse LU of the oracle's matrix (scipy) -> sols[node, var]; test-side stand-in for a
    Krylov solve converged to machine precision"""
    import scipy.sparse.linalg as spla
    x = spla.spsolve(sysm.scipy().tocsc(), sysm.rhs)
    return gather_solution(mesh, sysm, x)


def max_threads():
    return int(lib().fso_max_threads())


def recover_resultants(mesh: Mesh, sols, nu, em, t, quirks=QUIRKS_REFERENCE):
    """membrane stresses and bending moments at the element centroids, local element axes
    (fso_recover_resultants; thesis doc/shellelements.tex:524 and :1394-1403) -> (n_elem, 6)"""
    out = np.zeros(6 * mesh.n_elem)
    xyz = np.ascontiguousarray(mesh.xyz, np.float64)
    s = np.ascontiguousarray(sols, np.float64).reshape(-1)
    assert s.size == 6 * mesh.n_nodes
    lib().fso_recover_resultants(_p(xyz), C.c_int64(mesh.n_elem), _p(mesh.etype), _p(mesh.eptr), _p(mesh.enodes),
                                 C.c_double(nu), C.c_double(em), C.c_double(t), C.c_int(quirks), _p(s), _p(out))
    return out.reshape(-1, 6)

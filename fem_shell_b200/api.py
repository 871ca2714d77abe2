"""ctypes binding of include/femshell_b200.h (no torch types, no numerical code here).

Mirrors the reference's call sequence (src/fem-shell/fem-shell.cpp):
    Mesh::read + DirichletBoundary + init  (fs.cpp:35-37,90-125)  -> FemShell.set_mesh
    forces <- <mesh>_f                     (fs.cpp:44-67)         -> FemShell.set_nodal_loads
    equation_systems.solve()               (fs.cpp:138)           -> FemShell.assemble + FemShell.solve
    build_solution_vector                  (fs.cpp:140-141)       -> FemShell.solution
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# FEMSHELL_B200_LIB: lab builds of the same library (tools/asm_variants.sh); the product is the in-tree file
LIB_PATH = os.environ.get("FEMSHELL_B200_LIB") or os.path.join(HERE, "libfemshell_b200.so")

TRI3, QUAD4 = 3, 5
DOF_FIRST_ENCOUNTER, DOF_NODE_ID = 0, 1
PC_NONE, PC_JACOBI, PC_BJACOBI6, PC_MLRBM = 0, 1, 2, 3
NORM_UNPRECONDITIONED, NORM_PRECONDITIONED = 0, 1
QUIRKS_REFERENCE = 3
COMM_AUTO, COMM_NCCL, COMM_PEER = 0, 1, 2   # how the CG iteration talks between GPUs (fs_peer.cuh)
SPMV_AUTO, SPMV_FULL = 0, 1       # AUTO: iterate on the zero-compacted copy when the blocks share a planar pattern
ASM_COLORED, ASM_GATHER = 0, 1   # gather is the default; it falls back to coloured if a row is too dense
FS_OK, FS_ERR_ARG, FS_ERR_CUDA, FS_ERR_STATE, FS_ERR_NOT_CONVERGED, FS_ERR_BREAKDOWN, FS_ERR_COMM, FS_ERR_IO = 0, -1, -2, -3, -4, -5, -6, -7

# every symbol include/femshell_b200.h declares (checked by tests/test_abi.py)
EXPORTED_SYMBOLS = [
    "fs_create", "fs_destroy", "fs_last_error", "fs_get_stream", "fs_dist_unique_id", "fs_dist_init", "fs_set_comm_mode", "fs_get_comm_mode", "fs_get_comm_stats",
    "fs_set_material", "fs_set_quirks", "fs_set_dof_order", "fs_set_assembly_mode", "fs_set_spmv_format", "fs_get_spmv_format", "fs_set_mesh",
    "fs_set_nodal_loads", "fs_set_interface_loads", "fs_build_rhs", "fs_assemble", "fs_get_assembly_path", "fs_solve",
    "fs_get_solution", "fs_get_solution_owned", "fs_recover_resultants", "fs_solve_host", "fs_interface_nodes", "fs_step", "fs_commit_step", "fs_get_sizes",
    "fs_export_dof_order", "fs_export_csr", "fs_export_rhs", "fs_debug_element_matrices", "fs_spmv_host",
    "fs_bench_spmv", "fs_bench_fp64_peak", "fs_bench_contraction", "fs_bench_launch_chain", "fs_set_ml_options", "fs_get_ml_info", "fs_get_ml_dist_levels", "fs_get_ml_compact_levels", "fs_get_ml_profile", "fs_debug_ml_level", "fs_apply_mlrbm_host", "fs_partition_plan", "fs_gather_plan", "fs_meshgen", "fs_read_xda", "fs_read_mesh", "fs_write_xdr", "fs_read_forces", "fs_write_xda",
]


class FemShellError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("femshell_b200 error %d: %s" % (code, msg))
        self.code = code


class _Opts(C.Structure):
    _fields_ = [("rtol", C.c_double), ("max_its", C.c_int64), ("pc", C.c_int), ("norm_type", C.c_int),
                ("warm_start", C.c_int), ("check_every", C.c_int)]


class _Info(C.Structure):
    _fields_ = [("iterations", C.c_int64), ("rel_residual", C.c_double), ("status", C.c_int),
                ("solve_ms", C.c_float), ("spmv_ms", C.c_float)]


@dataclass
class SolveInfo:
    iterations: int
    rel_residual: float
    status: int
    solve_ms: float
    spmv_ms: float = 0.0


_lib = None


def load_library():
    """dlopen the CUDA library; raises if it has not been built (no fallback of any kind)"""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FemShellError(FS_ERR_STATE, "%s not built: run `python -c 'import __graft_entry__ as g; g.build()'`" % LIB_PATH)
        _lib = C.CDLL(LIB_PATH)
        _lib.fs_last_error.restype = C.c_char_p
        _lib.fs_get_stream.restype = C.c_void_p
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


# ----------------------------------------------------------------------------------------------
# host-side formats / generator (fs_meshio.cpp)
# ----------------------------------------------------------------------------------------------
def meshgen(kind, nx, ny, min_x, min_y, max_x, max_y, bcids, factor, loading, ul_lr, dead_axis="z"):
    """in-memory meshGen (src/meshgen/main_all.cpp); returns dict(xyz, etype, eptr, enodes, bc, forces)"""
    lib = load_library()
    nn, ne, nb = C.c_int64(), C.c_int64(), C.c_int64()
    bcs = (C.c_int * 4)(*[int(b) for b in bcids])
    args = [C.c_char(kind.encode()), C.c_int(nx), C.c_int(ny), C.c_double(min_x), C.c_double(min_y), C.c_double(max_x),
            C.c_double(max_y), bcs, C.c_double(factor), C.c_int(loading), C.c_int(1 if ul_lr else 0), C.c_char(dead_axis.encode())]
    rc = lib.fs_meshgen(*args, C.byref(nn), C.byref(ne), C.byref(nb), None, None, None, None, None, None)
    if rc:
        raise FemShellError(rc, "fs_meshgen: bad arguments")
    nen = 3 if kind.lower() == "t" else 4
    xyz = np.empty((nn.value, 3)); etype = np.empty(ne.value, np.int32); eptr = np.empty(ne.value + 1, np.int64)
    enodes = np.empty(ne.value * nen, np.int32); bc = np.empty((nb.value, 3), np.int32); F = np.empty((nn.value, 6))
    rc = lib.fs_meshgen(*args, C.byref(nn), C.byref(ne), C.byref(nb), _p(xyz), _p(etype), _p(eptr), _p(enodes), _p(bc), _p(F))
    if rc:
        raise FemShellError(rc, "fs_meshgen failed")
    return dict(xyz=xyz, etype=etype, eptr=eptr, enodes=enodes, bc=bc, forces=F)


def gather_plan(etype, eptr, enodes, n_nodes, mask=None, warp_vals=2816):
    """host-only schedule of the row-gather assembly pass (fs_gather_plan); None when the mesh needs the coloured pass"""
    lib = load_library()
    etype, eptr, enodes = _i32(etype), _i64(eptr), _i32(enodes)
    mk = None if mask is None else np.ascontiguousarray(mask, np.uint8)
    sizes = np.zeros(3, np.int64)
    args = [C.c_int64(n_nodes), C.c_int64(etype.size), _p(etype), _p(eptr), _p(enodes), _p(mk), C.c_int(warp_vals)]
    if lib.fs_gather_plan(*args, _p(sizes), None, None, None, None, None):
        raise FemShellError(-1, "fs_gather_plan: bad arguments")
    nptr = np.empty(n_nodes + 1, np.int32); nadj = np.empty(sizes[2], np.int32)
    if not sizes[1]:
        return None
    chunks = np.empty((sizes[0], 4), np.int64); info = np.empty((sizes[0] * 32, 4), np.int32); nodes = np.empty((sizes[0] * 32, 4), np.int32)
    if lib.fs_gather_plan(*args, _p(sizes), _p(chunks), _p(info), _p(nodes), _p(nptr), _p(nadj)):
        raise FemShellError(-1, "fs_gather_plan failed")
    return dict(chunks=chunks, info=info, nodes=nodes, nptr=nptr, nadj=nadj)


def partition_plan(eptr, enodes, n_nodes, rank, world, dof_mode=DOF_FIRST_ENCOUNTER):
    """host-only node-block partition plan of one rank (fs_partition_plan)"""
    lib = load_library()
    eptr, enodes = _i64(eptr), _i32(enodes)
    sizes = np.zeros(8, np.int64)
    args = [C.c_int64(n_nodes), C.c_int64(eptr.size - 1), _p(eptr), _p(enodes), C.c_int(dof_mode), C.c_int(rank), C.c_int(world)]
    rc = lib.fs_partition_plan(*args, _p(sizes), None, None, None, None)
    if rc:
        raise FemShellError(rc, "fs_partition_plan: bad arguments")
    l2g = np.empty(sizes[4], np.int32); le = np.empty(sizes[5], np.int32); si = np.empty(sizes[6], np.int32)
    pt = np.empty((sizes[7], 5), np.int64)
    rc = lib.fs_partition_plan(*args, _p(sizes), _p(l2g), _p(le), _p(si), _p(pt))
    if rc:
        raise FemShellError(rc, "fs_partition_plan failed")
    return dict(n_global=int(sizes[0]), own_begin=int(sizes[1]), own_end=int(sizes[2]), own_lo=int(sizes[3]),
                local_to_global=l2g, loc_elems=le, send_idx=si,
                peers=[dict(rank=int(r[0]), send_count=int(r[1]), send_off=int(r[2]), recv_count=int(r[3]), recv_off=int(r[4])) for r in pt])


def _read_with(fn_name, path):
    lib = load_library()
    fn = getattr(lib, fn_name)
    nn, ne, nen, nb = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
    rc = fn(path.encode(), C.byref(nn), C.byref(ne), C.byref(nen), C.byref(nb), None, None, None, None, None)
    if rc:
        raise FemShellError(rc, "%s(%s)" % (fn_name, path))
    xyz = np.empty((nn.value, 3)); etype = np.empty(ne.value, np.int32); eptr = np.empty(ne.value + 1, np.int64)
    enodes = np.empty(nen.value, np.int32); bc = np.empty((nb.value, 3), np.int32)
    rc = fn(path.encode(), C.byref(nn), C.byref(ne), C.byref(nen), C.byref(nb), _p(xyz), _p(etype), _p(eptr), _p(enodes), _p(bc))
    if rc:
        raise FemShellError(rc, "%s(%s)" % (fn_name, path))
    return dict(xyz=xyz, etype=etype, eptr=eptr, enodes=enodes, bc=bc)


def read_xda(path):
    return _read_with("fs_read_xda", path)


def read_mesh(path):
    """mesh.read() of fs.cpp:37: *.msh (Gmsh 2.x ASCII), *.xdr (binary twin of XDA) or XDA, by extension"""
    return _read_with("fs_read_mesh", path)


def write_xdr(path, xyz, etype, eptr, enodes, bc):
    xyz, etype, eptr, enodes, bc = _f64(xyz), _i32(etype), _i64(eptr), _i32(enodes), _i32(bc).reshape(-1, 3)
    rc = load_library().fs_write_xdr(path.encode(), C.c_int64(xyz.shape[0]), _p(xyz), C.c_int64(etype.size), _p(etype), _p(eptr), _p(enodes),
                                     C.c_int64(bc.shape[0]), _p(bc))
    if rc:
        raise FemShellError(rc, "fs_write_xdr(%s)" % path)


def read_forces(path, n_nodes):
    F = np.zeros((n_nodes, 6))
    rc = load_library().fs_read_forces(path.encode(), C.c_int64(n_nodes), _p(F))
    if rc:
        raise FemShellError(rc, "fs_read_forces(%s)" % path)
    return F


def write_xda(path, xyz, etype, eptr, enodes, bc):
    xyz, etype, eptr, enodes, bc = _f64(xyz), _i32(etype), _i64(eptr), _i32(enodes), _i32(bc).reshape(-1, 3)
    rc = load_library().fs_write_xda(path.encode(), C.c_int64(xyz.shape[0]), _p(xyz), C.c_int64(etype.size), _p(etype),
                                     _p(eptr), _p(enodes), C.c_int64(bc.shape[0]), _p(bc))
    if rc:
        raise FemShellError(rc, "fs_write_xda(%s)" % path)


# ----------------------------------------------------------------------------------------------
# the solver context
# ----------------------------------------------------------------------------------------------
class FemShell:
    """One GPU, one stream.  Method names follow the C ABI one to one."""

    def __init__(self, device=0, rank=0, world=1, nccl_id=None, comm=COMM_AUTO):
        self.lib = load_library()
        self.ctx = C.c_void_p()
        rc = self.lib.fs_create(C.byref(self.ctx), C.c_int(device))
        if rc:
            raise FemShellError(rc, "fs_create(device=%d) failed: no usable CUDA device (there is no CPU fallback)" % device)
        self.n_nodes = self.n_elem = 0
        self.world, self.rank = world, rank
        if world > 1:
            buf = (C.c_uint8 * 128).from_buffer_copy(bytes(nccl_id))
            self._ck(self.lib.fs_dist_init(self.ctx, C.c_int(rank), C.c_int(world), buf))
            self._ck(self.lib.fs_set_comm_mode(self.ctx, C.c_int(comm)))

    def comm_mode(self) -> int:
        m = C.c_int()
        self._ck(self.lib.fs_get_comm_mode(self.ctx, C.byref(m)))
        return m.value

    def comm_stats(self, reset=True):
        """microseconds block 0 of the CG kernels waited for the other GPUs (peer path) and the number of waits"""
        out = (C.c_double * 9)()
        self._ck(self.lib.fs_get_comm_stats(self.ctx, out, C.c_int(1 if reset else 0)))
        return {"halo_wait_us": out[0], "pq_wait_us": out[1], "rz_wait_us": out[2], "halo_waits": int(out[3]), "pq_waits": int(out[4]), "rz_waits": int(out[5]),
                "spmv_us": out[6], "update_us": out[7], "direction_us": out[8]}

    @staticmethod
    def unique_id() -> bytes:
        buf = (C.c_uint8 * 128)()
        rc = load_library().fs_dist_unique_id(buf)
        if rc:
            raise FemShellError(rc, "ncclGetUniqueId failed")
        return bytes(buf)

    def _ck(self, rc, allow=()):
        if rc != 0 and rc not in allow:
            raise FemShellError(rc, self.lib.fs_last_error(self.ctx).decode())
        return rc

    def close(self):
        if self.ctx:
            self.lib.fs_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def stream(self):
        return self.lib.fs_get_stream(self.ctx)

    # ---- inputs ----
    def set_material(self, nu, E, t):
        self._ck(self.lib.fs_set_material(self.ctx, C.c_double(nu), C.c_double(E), C.c_double(t)))

    def set_quirks(self, q):
        self._ck(self.lib.fs_set_quirks(self.ctx, C.c_int(q)))

    def set_dof_order(self, mode):
        self._ck(self.lib.fs_set_dof_order(self.ctx, C.c_int(mode)))

    def set_assembly_mode(self, mode):
        self._ck(self.lib.fs_set_assembly_mode(self.ctx, C.c_int(mode)))

    def set_spmv_format(self, mode):
        self._ck(self.lib.fs_set_spmv_format(self.ctx, C.c_int(mode)))

    def spmv_format(self):
        """{'nz_per_block': 36 | 14, 'matrix_bytes': streamed per SpMV, 'block_slots': .., 'pattern': 6x6 0/1 array}"""
        v = (C.c_int64 * 4)()
        self._ck(self.lib.fs_get_spmv_format(self.ctx, v))
        pat = np.array([[(v[3] >> (6 * a + b)) & 1 for b in range(6)] for a in range(6)], dtype=np.int32)
        return {"nz_per_block": int(v[0]), "matrix_bytes": int(v[1]), "block_slots": int(v[2]), "pattern": pat}

    def set_mesh(self, xyz, etype, eptr, enodes, bc):
        xyz, etype, eptr, enodes = _f64(xyz), _i32(etype), _i64(eptr), _i32(enodes)
        bc = _i32(bc).reshape(-1, 3)
        self.n_nodes, self.n_elem = xyz.shape[0], etype.size
        self._ck(self.lib.fs_set_mesh(self.ctx, C.c_int64(self.n_nodes), _p(xyz), C.c_int64(etype.size), _p(etype), _p(eptr),
                                      _p(enodes), C.c_int64(bc.shape[0]), _p(bc)))

    def set_nodal_loads(self, F):
        F = _f64(F)
        assert F.size == 6 * self.n_nodes
        self._ck(self.lib.fs_set_nodal_loads(self.ctx, _p(F)))

    def set_interface_loads(self, node_ids, dims, dead_axis, f):
        node_ids, f = _i32(node_ids), _f64(f)
        self._ck(self.lib.fs_set_interface_loads(self.ctx, C.c_int64(node_ids.size), _p(node_ids), C.c_int(dims),
                                                 C.c_char(dead_axis.encode()), _p(f)))

    def build_rhs(self, scale=1.0):
        self._ck(self.lib.fs_build_rhs(self.ctx, C.c_double(scale)))

    # ---- hot path ----
    def assemble(self) -> float:
        ms = C.c_float()
        self._ck(self.lib.fs_assemble(self.ctx, C.byref(ms)))
        return ms.value

    def assembly_path(self) -> str:
        p = C.c_int()
        self._ck(self.lib.fs_get_assembly_path(self.ctx, C.byref(p)))
        return ("k_assemble_colored", "k_assemble_gather", "k_assemble_slice")[p.value]

    @staticmethod
    def _opts(rtol, max_its, pc, norm_type, warm_start, check_every):
        return _Opts(rtol, max_its, pc, norm_type, 1 if warm_start else 0, check_every)

    def solve(self, rtol=1e-12, max_its=5000, pc=PC_JACOBI, norm_type=NORM_UNPRECONDITIONED, warm_start=True,
              check_every=0, allow_not_converged=False) -> SolveInfo:
        o, i = self._opts(rtol, max_its, pc, norm_type, warm_start, check_every), _Info()
        self._ck(self.lib.fs_solve(self.ctx, C.byref(o), C.byref(i)), allow=(FS_ERR_NOT_CONVERGED,) if allow_not_converged else ())
        return SolveInfo(i.iterations, i.rel_residual, i.status, i.solve_ms)

    def solution(self, out=None):
        sols = np.empty((self.n_nodes, 6)) if out is None else out
        self._ck(self.lib.fs_get_solution(self.ctx, _p(sols)))
        return sols

    def solution_owned(self, out=None, with_ids=True):
        """(node_ids, vals[n_own, 6]): this rank's rows of the solution in DOF order, no communication"""
        n = C.c_int64()
        self._ck(self.lib.fs_get_solution_owned(self.ctx, C.byref(n), None, None))
        ids = np.empty(n.value, np.int32) if with_ids else None
        vals = np.empty((n.value, 6)) if out is None else out
        self._ck(self.lib.fs_get_solution_owned(self.ctx, C.byref(n), _p(ids) if with_ids else None, _p(vals)))
        return ids, vals

    def recover_resultants(self):
        """(n_elem, 6): membrane stresses and bending moments at the element centroids, local element axes"""
        out = np.empty((self.n_elem, 6))
        self._ck(self.lib.fs_recover_resultants(self.ctx, _p(out)))
        return out

    def solve_host(self, F, sols, reassemble=False, rtol=1e-12, max_its=5000, pc=PC_JACOBI, norm_type=NORM_UNPRECONDITIONED,
                   warm_start=True, check_every=0, allow_not_converged=False) -> SolveInfo:
        """plugin-style call with host buffers: loads in, displacements out (fsp.cpp:271-274)"""
        o, i = self._opts(rtol, max_its, pc, norm_type, warm_start, check_every), _Info()
        self._ck(self.lib.fs_solve_host(self.ctx, _p(F), C.c_int(1 if reassemble else 0), C.byref(o), _p(sols) if sols is not None else None, C.byref(i)),
                 allow=(FS_ERR_NOT_CONVERGED,) if allow_not_converged else ())
        return SolveInfo(i.iterations, i.rel_residual, i.status, i.solve_ms)

    # ---- coupled step ----
    def interface_nodes(self):
        n = C.c_int64()
        self._ck(self.lib.fs_interface_nodes(self.ctx, C.byref(n), None))
        ids = np.empty(n.value, np.int32)
        self._ck(self.lib.fs_interface_nodes(self.ctx, C.byref(n), _p(ids)))
        return ids

    def step(self, dims, dead_axis, forces_in, rtol=1e-12, max_its=5000, pc=PC_JACOBI, norm_type=NORM_UNPRECONDITIONED,
             warm_start=True, allow_not_converged=False):
        f = _f64(forces_in)
        out = np.empty_like(f)
        o, i = self._opts(rtol, max_its, pc, norm_type, warm_start, 0), _Info()
        self._ck(self.lib.fs_step(self.ctx, C.c_int(dims), C.c_char(dead_axis.encode()), _p(f), C.byref(o), _p(out), C.byref(i)),
                 allow=(FS_ERR_NOT_CONVERGED,) if allow_not_converged else ())
        return out, SolveInfo(i.iterations, i.rel_residual, i.status, i.solve_ms)

    def commit_step(self, dims, dead_axis):
        self._ck(self.lib.fs_commit_step(self.ctx, C.c_int(dims), C.c_char(dead_axis.encode())))

    # ---- inspection ----
    def sizes(self):
        a = [C.c_int64() for _ in range(5)]
        self._ck(self.lib.fs_get_sizes(self.ctx, *[C.byref(v) for v in a]))
        return dict(n_dofnodes=a[0].value, n_blocks=a[1].value, n_colors=a[2].value, own_begin=a[3].value, own_end=a[4].value)

    def dof_order(self):
        d = np.empty(self.n_nodes, np.int32)
        self._ck(self.lib.fs_export_dof_order(self.ctx, _p(d)))
        return d

    def export_csr(self, with_cols=True, with_vals=True):
        s = self.sizes()
        n_own = s["own_end"] - s["own_begin"]
        rowptr = np.empty(6 * n_own + 1, np.int64)
        colidx = np.empty(36 * s["n_blocks"], np.int32) if with_cols else None
        vals = np.empty(36 * s["n_blocks"]) if with_vals else None
        self._ck(self.lib.fs_export_csr(self.ctx, _p(rowptr), _p(colidx), _p(vals)))
        return rowptr, colidx, vals

    def export_rhs(self):
        s = self.sizes()
        b = np.empty(6 * (s["own_end"] - s["own_begin"]))
        self._ck(self.lib.fs_export_rhs(self.ctx, _p(b)))
        return b

    def element_matrices(self, n_elem):
        out = np.zeros((n_elem, 576))
        self._ck(self.lib.fs_debug_element_matrices(self.ctx, _p(out)))
        return out

    def spmv(self, x):
        x = _f64(x)
        y = np.empty_like(x)
        self._ck(self.lib.fs_spmv_host(self.ctx, _p(x), _p(y)))
        return y

    def set_ml_options(self, max_points=1 << 22, dense_points=400, gamma=2):
        self._ck(self.lib.fs_set_ml_options(self.ctx, C.c_int64(max_points), C.c_int(dense_points), C.c_int(gamma)))

    def ml_info(self):
        lv = C.c_int64(0)
        cells = (C.c_int64 * 42)()
        w = (C.c_double * 15)()
        ms = C.c_double(0.0)
        self._ck(self.lib.fs_get_ml_info(self.ctx, C.byref(lv), cells, w, C.byref(ms)))
        n = int(lv.value)
        nd = C.c_int64(0)
        self._ck(self.lib.fs_get_ml_dist_levels(self.ctx, C.byref(nd)))
        nc = C.c_int64(0)
        self._ck(self.lib.fs_get_ml_compact_levels(self.ctx, C.byref(nc)))
        return {"levels": n, "distributed_levels": int(nd.value), "compact_levels": int(nc.value), "cells": [tuple(int(cells[3 * l + d]) for d in range(3)) for l in range(n)],
                "lambda": [float(w[i]) for i in range(n + 1)], "setup_ms": float(ms.value)}

    def bench_launch_chain(self, links=200, n=65536, reps=20):
        out = (C.c_double * 2)()
        self._ck(self.lib.fs_bench_launch_chain(self.ctx, C.c_int(links), C.c_int(n), C.c_int(reps), out))
        return {"plain_us_per_kernel": out[0], "pdl_us_per_kernel": out[1]}

    def ml_profile(self, reset=True):
        """FS_ML_PROFILE=1 runs: per-iteration stage times (ms) of the multilevel-preconditioned CG"""
        ms = (C.c_double * 8)()
        n = C.c_int64(0)
        self._ck(self.lib.fs_get_ml_profile(self.ctx, ms, C.byref(n), C.c_int(1 if reset else 0)))
        k = max(1, int(n.value))
        names = ("halo_spmv_update", "presmooth_restrict", "lattice_cycle", "prolong", "postsmooth_spmv_rz", "allreduce_direction",
                 "lattice_level2_visit1", "lattice_level2_visit2")
        return {"iterations": int(n.value), **{nm: float(ms[i]) / k for i, nm in enumerate(names)}}

    def ml_level(self, level, what):
        n = C.c_int64(0)
        self._ck(self.lib.fs_debug_ml_level(self.ctx, C.c_int(level), C.c_int(what), None, C.c_int64(0), C.byref(n)))
        out = np.empty(n.value)
        self._ck(self.lib.fs_debug_ml_level(self.ctx, C.c_int(level), C.c_int(what), _p(out), C.c_int64(out.size), C.byref(n)))
        return out

    def apply_mlrbm(self, r):
        r = _f64(r).ravel()
        z = np.empty_like(r)
        self._ck(self.lib.fs_apply_mlrbm_host(self.ctx, _p(r), _p(z)))
        return z

    def bench_fp64_peak(self) -> float:
        t = C.c_double()
        self._ck(self.lib.fs_bench_fp64_peak(self.ctx, C.byref(t)))
        return t.value

    def bench_contraction(self, n_elem=1 << 20, reps=10):
        """DMMA vs FMA-pipe micro-benchmark of the plate contraction (fs_bench_contraction)"""
        out = (C.c_double * 6)()
        self._ck(self.lib.fs_bench_contraction(self.ctx, C.c_int64(n_elem), C.c_int(reps), out))
        return {"fma_ms": out[0], "dmma_ms": out[1], "max_rel_diff": out[2], "dmma_peak_tflops": out[3],
                "fma_useful_tflops": out[4], "dmma_useful_tflops": out[5]}

    def bench_spmv(self, reps=20) -> float:
        i = _Info()
        self._ck(self.lib.fs_bench_spmv(self.ctx, C.c_int(reps), C.byref(i)))
        self.last_spmv_ms_on_p = i.solve_ms
        return i.spmv_ms

"""torchrun worker for the multi-GPU parity test (tests/test_gpu_multi.py): every rank drives one GPU
through the C ABI; rank-local CSR rows and the distributed solve are checked against the CPU oracle
and against a single-GPU context."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch
import torch.distributed as dist

import fem_shell_b200 as fsb
import meshes
from oracle import fso


def distributed_lattice_case(rank, world, lr):
    """multilevel preconditioner with DISTRIBUTED lattice levels (fs_mlpc.cu): a plate large enough for two lattice
    levels above the dense one; slabs of cells per rank, halo slabs exchanged, boundary slabs of the restrictions
    added at their owner.  Same iteration count as one GPU, same displacements as the oracle."""
    os.environ["FS_ML_DIST_MIN_CELLS"] = "64"
    for kind, nx, ny in (("t", 200, 151), ("q", 181, 230)):
        m = fsb.meshgen(kind, nx, ny, 0, 0, 10, 10.0 * ny / nx, (1, 1, 1, 1), 300.0, 2, 1)
        om = fso.Mesh(np.asarray(m["xyz"], float), m["etype"], m["eptr"], m["enodes"], m["bc"])
        ref = fso.assemble(om, m["forces"], 0.3, 1e7, 0.5)
        s1 = fsb.FemShell(device=lr)
        s1.set_material(0.3, 1e7, 0.5)
        s1.set_mesh(m["xyz"], m["etype"], m["eptr"], m["enodes"], m["bc"])
        s1.set_nodal_loads(m["forces"])
        s1.assemble()
        i1 = s1.solve(rtol=1e-10, max_its=3000, pc=fsb.PC_MLRBM, warm_start=False)
        u1 = s1.solution()
        x1 = np.zeros(6 * ref.n_dofnodes); x1.reshape(-1, 6)[ref.dofnode] = u1
        r1 = np.linalg.norm(ref.rhs - fso.spmv(ref, x1)) / np.linalg.norm(ref.rhs)
        s1.close()
        for comm in (fsb.COMM_PEER, fsb.COMM_NCCL):
            ids = [fsb.FemShell.unique_id() if rank == 0 else None]
            dist.broadcast_object_list(ids, src=0)
            s = fsb.FemShell(device=lr, rank=rank, world=world, nccl_id=ids[0], comm=comm)
            s.set_material(0.3, 1e7, 0.5)
            s.set_mesh(m["xyz"], m["etype"], m["eptr"], m["enodes"], m["bc"])
            s.set_nodal_loads(m["forces"])
            s.assemble()
            info = s.solve(rtol=1e-10, max_its=3000, pc=fsb.PC_MLRBM, warm_start=False)
            mi = s.ml_info()
            assert mi["distributed_levels"] >= 1, mi
            u = s.solution()
            xu = np.zeros(6 * ref.n_dofnodes); xu.reshape(-1, 6)[ref.dofnode] = u
            ru = np.linalg.norm(ref.rhs - fso.spmv(ref, xu)) / np.linalg.norm(ref.rhs)
            assert abs(info.iterations - i1.iterations) <= 2, (kind, comm, info.iterations, i1.iterations)
            assert np.linalg.norm(u - u1) <= 1e-8 * np.linalg.norm(u1), (kind, comm, np.linalg.norm(u - u1) / np.linalg.norm(u1))
            # a re-assembly with another material rebuilds the distributed stencils (set-up path a second time)
            s.set_material(0.25, 2e7, 0.4)
            s.assemble()
            info2 = s.solve(rtol=1e-10, max_its=3000, pc=fsb.PC_MLRBM, warm_start=False)
            ref2 = fso.assemble(om, m["forces"], 0.25, 2e7, 0.4)
            x2 = np.zeros(6 * ref.n_dofnodes); x2.reshape(-1, 6)[ref.dofnode] = s.solution()
            r2 = np.linalg.norm(ref2.rhs - fso.spmv(ref2, x2)) / np.linalg.norm(ref2.rhs)
            assert r2 <= 100 * max(r1, 1e-10), (kind, comm, r2, r1)
            s.close()
            dist.barrier()
            if rank == 0:
                print("dist lattice ok %s world=%d comm=%d: %d distributed of %d levels %s, iterations %d (single %d), oracle residual %.2e (single %.2e), second material %d its %.2e"
                      % (kind, world, comm, mi["distributed_levels"], mi["levels"], mi["cells"], info.iterations, i1.iterations,
                         ru, r1, info2.iterations, r2), flush=True)
    del os.environ["FS_ML_DIST_MIN_CELLS"]


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    cases = {
        "tri": (fsb.meshgen("t", 40, 33, 0, 0, 10, 10, (1, 1, 1, 1), 300.0, 2, 1), 0.3, 1e7, 0.5),
        "quad": (fsb.meshgen("q", 31, 47, 0, 0, 10, 10, (0, 1, 0, 1), 300.0, 2, 1), 0.3, 1e7, 0.5),
        "mixed": (meshes.folded_cantilever(nx=24, ny=10, skew=0.3), 0.3, 1e4, 0.25),
    }
    for name, (m, nu, E, t) in cases.items():
        om = fso.Mesh(np.asarray(m["xyz"], float), m["etype"], m["eptr"], m["enodes"], m["bc"])
        ref = fso.assemble(om, m["forces"], nu, E, t)
        uo = fso.direct_solve(om, ref)
        # single-GPU context on the same mesh: iteration counts must agree (+-1%)
        s1 = fsb.FemShell(device=lr)
        s1.set_material(nu, E, t)
        s1.set_mesh(m["xyz"], m["etype"], m["eptr"], m["enodes"], m["bc"])
        s1.set_nodal_loads(m["forces"])
        s1.assemble()
        i1 = s1.solve(rtol=1e-12, max_its=400000, pc=fsb.PC_BJACOBI6, warm_start=False)
        u1 = s1.solution()
        m1 = s1.solve(rtol=1e-12, max_its=3000, pc=fsb.PC_MLRBM, warm_start=False)   # multilevel cycle, single GPU
        assert np.linalg.norm(s1.solution() - uo) <= 1e-8 * np.linalg.norm(uo), name
        s1.close()
        its = {}
        # NVLink peer windows (kernels push halos / partial sums themselves) and the NCCL path
        for comm in (fsb.COMM_PEER, fsb.COMM_NCCL):
            ids = [fsb.FemShell.unique_id() if rank == 0 else None]   # an ncclUniqueId serves one communicator
            dist.broadcast_object_list(ids, src=0)
            s = fsb.FemShell(device=lr, rank=rank, world=world, nccl_id=ids[0], comm=comm)
            s.set_assembly_mode(fsb.ASM_GATHER if name != "quad" else fsb.ASM_COLORED)
            s.set_material(nu, E, t)
            s.set_mesh(m["xyz"], m["etype"], m["eptr"], m["enodes"], m["bc"])
            assert s.comm_mode() == comm, (name, comm, s.comm_mode())
            s.set_nodal_loads(m["forces"])
            s.assemble()
            sz = s.sizes()
            ob, oe = sz["own_begin"], sz["own_end"]
            rowptr, colidx, vals = s.export_csr()
            rr, rc, rv = ref.csr()
            lo, hi = rr[6 * ob], rr[6 * oe]
            assert np.array_equal(rowptr, rr[6 * ob:6 * oe + 1] - lo), name
            assert np.array_equal(colidx, rc[lo:hi]), name
            assert np.abs(vals - rv[lo:hi]).max() <= 1e-12 * np.abs(rv).max(), name
            assert np.array_equal(s.export_rhs(), ref.rhs[6 * ob:6 * oe]), name
            info = s.solve(rtol=1e-12, max_its=400000, pc=fsb.PC_BJACOBI6, warm_start=False)
            u = s.solution()
            err = np.linalg.norm(u - uo) / np.linalg.norm(uo)
            assert err <= 1e-8, (name, comm, err)
            assert abs(i1.iterations - info.iterations) <= max(3, i1.iterations // 50), (name, comm, i1.iterations, info.iterations)
            assert np.linalg.norm(u - u1) <= 1e-8 * np.linalg.norm(u1)
            its[comm] = info.iterations
            sols_by_comm = locals().setdefault("sols_by_comm", {})
            sols_by_comm[(name, comm)] = u
            ids, own = s.solution_owned()          # this rank's rows, no communication
            assert ids.size == oe - ob and np.array_equal(own, u[ids]), (name, comm, "owned rows")
            # stress resultants: elements of the cut rows read halo displacements; every rank gets all rows
            res = s.recover_resultants()
            ro = fso.recover_resultants(om, u, nu, E, t)
            assert np.abs(res - ro).max() <= 1e-10 * np.abs(ro).max(), (name, comm, "resultants")
            # multilevel preconditioner across ranks: lattice levels replicated, restricted residual all-reduced
            mi = s.solve(rtol=1e-12, max_its=3000, pc=fsb.PC_MLRBM, warm_start=False)
            um = s.solution()
            assert np.linalg.norm(um - uo) <= 1e-8 * np.linalg.norm(uo), (name, comm, "mlrbm")
            assert abs(mi.iterations - m1.iterations) <= 2, (name, comm, mi.iterations, m1.iterations)
            # a second load case on the same matrix (stamps / mailboxes carry over between solves), Jacobi this time
            s.build_rhs(-2.5)
            s.solve(rtol=1e-12, max_its=400000, pc=fsb.PC_JACOBI, warm_start=True)
            u2 = s.solution()
            assert np.linalg.norm(u2 + 2.5 * uo) <= 1e-8 * np.linalg.norm(2.5 * uo), (name, comm)
            s.close()
            dist.barrier()
        assert abs(its[fsb.COMM_PEER] - its[fsb.COMM_NCCL]) <= max(3, i1.iterations // 50), (name, its)
        # thousands of iterations through the peer windows (halo pushed by k_direction, consumed after the interior
        # slices of the SpMV) and through NCCL give the same displacements: a stale halo read would not
        du = np.linalg.norm(sols_by_comm[(name, fsb.COMM_PEER)] - sols_by_comm[(name, fsb.COMM_NCCL)]) / np.linalg.norm(uo)
        assert du <= 1e-10, (name, du)
        if rank == 0:
            print("dist ok %-6s world=%d iterations peer %d nccl %d (single %d) multilevel %d (single %d) err %.2e"
                  % (name, world, its[fsb.COMM_PEER], its[fsb.COMM_NCCL], i1.iterations, mi.iterations, m1.iterations, err), flush=True)
    distributed_lattice_case(rank, world, lr)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

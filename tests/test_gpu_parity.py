"""GPU parity tests: the CUDA path (through the C ABI, via ctypes) against the CPU oracle on the
same inputs.  Bars (BASELINE.json north_star):
  * CSR sparsity pattern and DOF ordering       bit-exact
  * assembled stiffness entries                 |d| <= 1e-12 * max|6x6 block|   (FP64, summation order only)
  * displacements                               <= 1e-8 relative L2
"""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_ref_mesh
import meshes

pytestmark = pytest.mark.gpu

GOLD = json.load(open(os.path.join(GOLDEN, "thesis_goldens.json")))


@pytest.fixture(scope="module")
def fsb():
    import fem_shell_b200 as fsb
    return fsb


def as_fso_mesh(fso, m):
    return fso.Mesh(np.asarray(m["xyz"], float), m["etype"], m["eptr"], m["enodes"], m["bc"])


def gpu_system(fsb, m, nu, E, t, dof=0, quirks=3, loads=None, asm=0):
    s = fsb.FemShell()
    s.set_assembly_mode(asm)
    s.set_material(nu, E, t)
    s.set_quirks(quirks)
    s.set_dof_order(dof)
    s.set_mesh(m["xyz"], m["etype"], m["eptr"], m["enodes"], m["bc"])
    if loads is not None:
        s.set_nodal_loads(loads)
    s.assemble()
    return s


def block_scaled_error(vals, ref, nptr):
    """max over 6x6 blocks of |d| / max|ref block|  (rows of node p are contiguous: 6 x 6deg)"""
    worst = 0.0
    for p in range(nptr.size - 1):
        deg = nptr[p + 1] - nptr[p]
        a = vals[36 * nptr[p]:36 * nptr[p + 1]].reshape(6, deg, 6)
        b = ref[36 * nptr[p]:36 * nptr[p + 1]].reshape(6, deg, 6)
        scale = np.abs(b).max(axis=(0, 2))
        scale[scale == 0] = 1.0
        worst = max(worst, (np.abs(a - b).max(axis=(0, 2)) / scale).max())
    return worst


# ---- element kernels -----------------------------------------------------------------------
@pytest.mark.parametrize("quirks", [3, 0])
def test_element_matrices_general_elements(fso, fsb, quirks):
    elems = meshes.random_elements(64)
    m = meshes.elements_as_mesh(elems)
    s = fsb.FemShell()
    s.set_material(0.3, 1e7, 0.05)
    s.set_quirks(quirks)
    s.set_mesh(m["xyz"], m["etype"], m["eptr"], m["enodes"], m["bc"])
    K = s.element_matrices(len(elems))
    for e, (et, P) in enumerate(elems):
        n6 = 18 if et == fso.TRI3 else 24
        ref = fso.element_stiffness(et, P, 0.3, 1e7, 0.05, quirks=quirks, layout=0)
        got = K[e, :n6 * n6].reshape(n6, n6)
        # badly shaped random elements: entries of the coupling blocks are differences of terms as
        # large as the diagonal blocks, so the summation-order noise scales with the element's
        # largest entry, not with the (cancelled) block -- SURVEY.md section 7 "1e-12 needs a scale"
        assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max(), (e, et, np.abs(got - ref).max() / np.abs(ref).max())


def test_element_matrices_meshgen_cells(fso, fsb):
    for kind, ul in (("q", 1), ("t", 1), ("t", 0)):
        m = fsb.meshgen(kind, 7, 5, 0, 0, 10, 3, (1, 1, 1, 1), 1.0, 2, ul)
        s = fsb.FemShell()
        s.set_material(0.3, 1e7, 0.5)
        s.set_mesh(m["xyz"], m["etype"], m["eptr"], m["enodes"], m["bc"])
        K = s.element_matrices(m["etype"].size)
        for e in range(m["etype"].size):
            ids = m["enodes"][m["eptr"][e]:m["eptr"][e + 1]]
            n6 = 6 * len(ids)
            ref = fso.element_stiffness(int(m["etype"][e]), m["xyz"][ids], 0.3, 1e7, 0.5)
            got = K[e, :n6 * n6].reshape(n6, n6)
            assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max()


# ---- pattern / values / rhs ----------------------------------------------------------------
CASES = {
    "c1_tri16": lambda fsb: (fsb.meshgen("t", 16, 16, 0, 0, 10, 10, (1, 1, 1, 1), 300.0, 2, 1), 0.3, 1e7, 0.5),
    "quad24": lambda fsb: (fsb.meshgen("q", 24, 17, 0, 0, 10, 7, (0, 1, 20, 21), 300.0, 2, 1), 0.3, 1e7, 0.5),
    "tri_ur": lambda fsb: (fsb.meshgen("t", 9, 13, -1, 0, 4, 6, (0, 0, -1, 1), 2.0, 1, 0, "y"), 0.25, 3e4, 1.0),
    "c5_folded": lambda fsb: (meshes.folded_cantilever(), 0.3, 1e4, 0.25),
    "c5_skewed": lambda fsb: (meshes.folded_cantilever(skew=0.35), 0.3, 1e4, 0.25),
    # unstructured: 14 triangles share the hub as their first local node (14 emit phases), mixed rings behind them
    "umbrella_mixed": lambda fsb: (meshes.umbrella(mixed=True, n_rings=5), 0.3, 1e4, 0.25),
    "umbrella_quads": lambda fsb: (meshes.umbrella(mixed=False, n_rings=4), 0.3, 1e4, 0.25),
    # 40 elements at one node do not fit a warp of the row-gather pass: the context falls back to the coloured pass
    "umbrella_hub40": lambda fsb: (meshes.umbrella(n_spokes=40, mixed=True, n_rings=3), 0.3, 1e4, 0.25),
}


@pytest.mark.parametrize("asm", [0, 1], ids=["colored", "gather"])
@pytest.mark.parametrize("dof", [0, 1])
@pytest.mark.parametrize("case", sorted(CASES))
def test_assembled_system_matches_oracle(fso, fsb, case, dof, asm):
    m, nu, E, t = CASES[case](fsb)
    om = as_fso_mesh(fso, m)
    ref = fso.assemble(om, m["forces"], nu, E, t, dof_mode=dof)
    s = gpu_system(fsb, m, nu, E, t, dof=dof, loads=m["forces"], asm=asm)
    assert np.array_equal(s.dof_order(), ref.dofnode)
    rowptr, colidx, vals = s.export_csr()
    rrow, rcol, rvals = ref.csr()
    assert np.array_equal(rowptr, rrow), "CSR row pointers differ"
    assert np.array_equal(colidx, rcol), "CSR column indices differ"
    assert block_scaled_error(vals, rvals, ref.nptr) <= 1e-12
    assert np.array_equal(s.export_rhs(), ref.rhs)
    # Dirichlet rows: exact integers on the diagonal, exact zeros elsewhere
    con = np.repeat(ref.mask[np.argsort(ref.dofnode)], 6) >> np.tile(np.arange(6), ref.n_dofnodes) & 1
    A = ref.scipy()
    A.data = vals.copy()
    d = A.diagonal()
    assert np.all(d[con == 1] == np.round(d[con == 1])) and np.all(d[con == 1] >= 1)
    assert abs(A[con == 1]).sum() == d[con == 1].sum()


def test_gather_and_colored_agree_at_scale(fsb):
    """both assembly strategies on a 300x300-cell mixed-size plate: same matrix to summation order"""
    for kind in ("q", "t"):
        m = fsb.meshgen(kind, 300, 200, 0, 0, 10, 7, (1, 0, 1, 0), 300.0, 2, 1)
        a = gpu_system(fsb, m, 0.3, 1e7, 0.5, asm=0).export_csr(with_cols=False)[2]
        b = gpu_system(fsb, m, 0.3, 1e7, 0.5, asm=1).export_csr(with_cols=False)[2]
        assert np.abs(a - b).max() <= 1e-12 * np.abs(a).max()
        # the gather pass is bitwise reproducible
        b2 = gpu_system(fsb, m, 0.3, 1e7, 0.5, asm=1).export_csr(with_cols=False)[2]
        assert np.array_equal(b, b2)


def test_spmv_matches_oracle(fso, fsb):
    m, nu, E, t = CASES["c5_skewed"](fsb)
    ref = fso.assemble(as_fso_mesh(fso, m), m["forces"], nu, E, t)
    s = gpu_system(fsb, m, nu, E, t, loads=m["forces"])
    x = np.random.default_rng(7).standard_normal(6 * ref.n_dofnodes)
    y, yr = s.spmv(x), fso.spmv(ref, x)
    assert np.abs(y - yr).max() <= 1e-12 * np.abs(yr).max()


def _rotate_mesh(m, R):
    m = dict(m)
    m["xyz"] = np.ascontiguousarray(np.asarray(m["xyz"], float) @ R.T)
    m["forces"] = np.ascontiguousarray(np.hstack([m["forces"][:, :3] @ R.T, m["forces"][:, 3:] @ R.T]))
    return m


@pytest.mark.parametrize("plane", ["xy", "xz", "yz"])
@pytest.mark.parametrize("kind", ["q", "t"])
def test_compacted_spmv_format_on_planar_shells(fso, fsb, kind, plane):
    """plates in a coordinate plane: the union pattern of the 6x6 blocks has 14 entries, the SpMV runs on the
    sliced-ELL copy and agrees with the oracle's CSR product and with the parity-format kernel"""
    m = fsb.meshgen(kind, 37, 21, 0, 0, 10, 6, (1, 0, 1, 20), 300.0, 2, 1)
    R = {"xy": np.eye(3), "xz": np.array([[1., 0, 0], [0, 0, -1], [0, 1, 0]]), "yz": np.array([[0., 0, 1], [1, 0, 0], [0, 1, 0]])}[plane]
    m = _rotate_mesh(m, R)
    ref = fso.assemble(as_fso_mesh(fso, m), m["forces"], 0.3, 1e7, 0.5)
    s = gpu_system(fsb, m, 0.3, 1e7, 0.5, loads=m["forces"])
    fmt = s.spmv_format()
    assert fmt["nz_per_block"] == 14 and fmt["pattern"].sum() == 14, fmt
    rp, ci, v = ref.csr()
    rows = np.repeat(np.arange(rp.size - 1), np.diff(rp))
    pat = np.zeros((6, 6), int)
    np.add.at(pat, (rows[v != 0] % 6, ci[v != 0] % 6), 1)
    assert np.array_equal(fmt["pattern"] > 0, pat > 0)
    x = np.random.default_rng(11).standard_normal(6 * ref.n_dofnodes)
    y, yr = s.spmv(x), fso.spmv(ref, x)
    assert np.abs(y - yr).max() <= 1e-13 * np.abs(yr).max()
    s.set_spmv_format(fsb.SPMV_FULL)
    assert s.spmv_format()["nz_per_block"] == 36
    y36 = s.spmv(x)
    assert np.abs(y - y36).max() <= 1e-13 * np.abs(yr).max()
    # the solve is the same on either copy
    i36 = s.solve(rtol=1e-10, max_its=200000, warm_start=False)
    u36 = s.solution()
    s.set_spmv_format(fsb.SPMV_AUTO)
    i14 = s.solve(rtol=1e-10, max_its=200000, warm_start=False)
    u14 = s.solution()
    assert abs(i14.iterations - i36.iterations) <= max(3, i36.iterations // 50)
    assert np.linalg.norm(u14 - u36) <= 1e-8 * np.linalg.norm(u36)
    # re-assembly with another material keeps the layout and refreshes the values
    s.set_material(0.2, 2e7, 0.4)
    s.assemble()
    ref2 = fso.assemble(as_fso_mesh(fso, m), m["forces"], 0.2, 2e7, 0.4)
    y2, y2r = s.spmv(x), fso.spmv(ref2, x)
    assert s.spmv_format()["nz_per_block"] == 14
    assert np.abs(y2 - y2r).max() <= 1e-13 * np.abs(y2r).max()


def test_dense_blocks_keep_the_parity_format(fso, fsb):
    m = meshes.folded_cantilever(skew=0.35)
    s = gpu_system(fsb, m, 0.3, 1e4, 0.25, loads=m["forces"])
    fmt = s.spmv_format()
    assert fmt["nz_per_block"] == 36 and fmt["pattern"].sum() == 36


# ---- solves ---------------------------------------------------------------------------------
def agree6(value, gold):
    import math
    ulp6 = 10.0 ** (math.floor(math.log10(abs(gold))) - 5)
    return abs(value - gold) <= 0.5001 * ulp6


@pytest.mark.parametrize("case", GOLD["shipped"], ids=[c["test"] for c in GOLD["shipped"]])
def test_thesis_goldens_on_gpu(fso, fsb, ref_meshes, case):
    mesh, F = load_ref_mesh(fso, ref_meshes, case["mesh"])
    m = dict(xyz=mesh.xyz, etype=mesh.etype, eptr=mesh.eptr, enodes=mesh.enodes, bc=mesh.bc)
    s = gpu_system(fsb, m, case["nu"], case["E"], case["t"], loads=F)
    info = s.solve(rtol=1e-12, max_its=200000, pc=fsb.PC_JACOBI)
    u = s.solution()
    for node, var, gold in case["checks"]:
        assert agree6(u[node, var], gold), (case["test"], node, var, u[node, var], gold)
    uo = fso.direct_solve(mesh, fso.assemble(mesh, F, case["nu"], case["E"], case["t"]))
    assert np.linalg.norm(u - uo) <= 1e-8 * np.linalg.norm(uo), info


@pytest.mark.parametrize("pc", [0, 1, 2])
@pytest.mark.parametrize("norm", [0, 1])
def test_pcg_same_iterations_as_oracle(fso, fsb, pc, norm):
    m, nu, E, t = CASES["c1_tri16"](fsb)
    ref = fso.assemble(as_fso_mesh(fso, m), m["forces"], nu, E, t)
    xo, its_o, rel_o = fso.pcg(ref, pc=pc, norm_type=norm, rtol=1e-8, max_its=100000)
    s = gpu_system(fsb, m, nu, E, t, loads=m["forces"])
    info = s.solve(rtol=1e-8, max_its=100000, pc=pc, norm_type=norm, warm_start=False)
    assert info.status == 0 and info.rel_residual <= 1e-8
    assert abs(info.iterations - its_o) <= max(3, its_o // 50), (info.iterations, its_o)
    u = s.solution()
    uo = fso.gather_solution(as_fso_mesh(fso, m), ref, xo)
    assert np.linalg.norm(u - uo) <= 1e-6 * np.linalg.norm(uo)


def test_mixed_folded_cantilever_solution(fso, fsb):
    for skew in (0.0, 0.35):
        m = meshes.folded_cantilever(skew=skew)
        om = as_fso_mesh(fso, m)
        uo = fso.direct_solve(om, fso.assemble(om, m["forces"], 0.3, 1e4, 0.25))
        s = gpu_system(fsb, m, 0.3, 1e4, 0.25, loads=m["forces"])
        s.solve(rtol=1e-13, max_its=400000, pc=fsb.PC_BJACOBI6)
        u = s.solution()
        assert np.linalg.norm(u - uo) <= 1e-8 * np.linalg.norm(uo)


@pytest.mark.gpu
@pytest.mark.parametrize("quirks", [3, 0])
def test_stress_resultants_match_oracle(fso, fsb, quirks):
    """fs_recover_resultants against the oracle's restatement on the SAME displacements:
    mixed skewed folded cantilever (both element types, rotated frames, quirks on and off) and a meshGen plate"""
    cases = [(meshes.folded_cantilever(skew=0.35), 0.3, 1e4, 0.25), (meshes.folded_cantilever(skew=0.0), 0.3, 1e4, 0.25),
             (fsb.meshgen("t", 16, 16, 0, 0, 10, 10, (1, 1, 1, 1), 300.0, 2, 1), 0.3, 1e7, 0.5),
             (fsb.meshgen("q", 12, 9, 0, 0, 10, 7, (0, 1, 0, 1), 300.0, 2, 1), 0.3, 1e7, 0.5)]
    for m, nu, E, t in cases:
        s = gpu_system(fsb, m, nu, E, t, quirks=quirks, loads=m["forces"])
        with pytest.raises(fsb.FemShellError):
            s.recover_resultants()          # no solution yet
        s.solve(rtol=1e-12, max_its=400000, pc=fsb.PC_BJACOBI6)
        u = s.solution()
        res = s.recover_resultants()
        ro = fso.recover_resultants(as_fso_mesh(fso, m), u, nu, E, t, quirks=quirks)
        assert res.shape == ro.shape
        for cols in (slice(0, 3), slice(3, 6)):   # bit-level agreement is not required: FMA contraction differs
            assert np.abs(res[:, cols] - ro[:, cols]).max() <= 1e-10 * np.abs(ro[:, cols]).max()
        s.close()


def test_owned_rows_of_the_solution(fso, fsb):
    """fs_get_solution_owned on one rank = the whole solution, addressed through the DOF order"""
    m, nu, E, t = CASES["c5_folded"](fsb)
    s = gpu_system(fsb, m, nu, E, t, loads=m["forces"])
    F = np.ascontiguousarray(m["forces"], float)
    info = s.solve_host(F, None, rtol=1e-12, max_its=400000, pc=fsb.PC_BJACOBI6, warm_start=False)   # sols = NULL: nothing gathered
    assert info.status == 0
    ids, vals = s.solution_owned()
    u = s.solution()
    assert sorted(ids.tolist()) == list(range(u.shape[0]))
    assert np.array_equal(vals, u[ids])
    assert np.array_equal(s.dof_order()[ids], np.arange(ids.size))   # row k = k-th node of the DOF order
    s.close()


def test_solve_host_takes_new_loads_with_every_reassembly(fso, fsb):
    """the plugin call of the coupling loop (fsp.cpp:271-274): new loads in, values pass, solve, displacements out --
    the load copy overlaps the values pass on a second stream, so stale or half-copied loads would show here"""
    m, nu, E, t = CASES["quad24"](fsb)
    om = as_fso_mesh(fso, m)
    s = gpu_system(fsb, m, nu, E, t)
    rng = np.random.default_rng(5)
    U = np.empty((om.n_nodes, 6))
    for k in range(4):
        F = np.ascontiguousarray(m["forces"] * (1.0 + k) + rng.standard_normal(m["forces"].shape) * (k % 2))
        info = s.solve_host(F, U, reassemble=(k != 2), rtol=1e-13, max_its=400000, pc=fsb.PC_BJACOBI6, warm_start=(k > 1))
        assert info.status == 0
        uo = fso.direct_solve(om, fso.assemble(om, F, nu, E, t))
        assert np.linalg.norm(U - uo) <= 1e-8 * np.linalg.norm(uo), k
    s.close()


def test_max_its_and_error_paths(fso, fsb):
    m, nu, E, t = CASES["quad24"](fsb)
    s = gpu_system(fsb, m, nu, E, t, loads=m["forces"])
    with pytest.raises(fsb.FemShellError) as ei:
        s.solve(rtol=1e-14, max_its=7, warm_start=False)
    assert ei.value.code == fsb.FS_ERR_NOT_CONVERGED
    info = s.solve(rtol=1e-14, max_its=7, warm_start=False, allow_not_converged=True)
    assert info.iterations == 7
    # zero rhs -> zero solution, zero iterations
    s.set_nodal_loads(np.zeros((s.n_nodes, 6)))
    info = s.solve(rtol=1e-8)
    assert info.iterations == 0 and not s.solution().any()
    s2 = fsb.FemShell()
    with pytest.raises(fsb.FemShellError):
        s2.assemble()          # no mesh
    bad = dict(m)
    bad["etype"] = m["etype"].copy()
    bad["etype"][0] = 9
    s2.set_material(nu, E, t)
    with pytest.raises(fsb.FemShellError):
        s2.set_mesh(bad["xyz"], bad["etype"], bad["eptr"], bad["enodes"], bad["bc"])


def test_repeated_rhs_and_warm_start(fso, fsb):
    """BASELINE config 4: fixed stiffness, changing pressure amplitude 1+sin(tau/25.01) (fluid_solver.cpp:192)"""
    m, nu, E, t = CASES["quad24"](fsb)
    om = as_fso_mesh(fso, m)
    ref = fso.assemble(om, m["forces"], nu, E, t)
    u1 = fso.direct_solve(om, ref)
    s = gpu_system(fsb, m, nu, E, t, loads=m["forces"])
    cold = s.solve(rtol=1e-10, max_its=100000, warm_start=False).iterations
    for tau in range(0, 100, 11):
        a = 1.0 + np.sin(tau / 25.01)
        s.build_rhs(a)
        info = s.solve(rtol=1e-10, max_its=100000, warm_start=True)
        assert info.iterations <= 1.25 * cold + 5, (info.iterations, cold)
        assert np.linalg.norm(s.solution() - a * u1) <= 1e-8 * np.linalg.norm(a * u1)


def test_coupled_step_contract(fso, fsb, ref_meshes):
    """fsp.cpp:257-374 on the reference's tower mesh: interface = ids 2/20/21, 2-D coupling, dead axis y... here z"""
    mesh, _ = load_ref_mesh(fso, ref_meshes, "bending_tower_tri_test")
    m = dict(xyz=mesh.xyz, etype=mesh.etype, eptr=mesh.eptr, enodes=mesh.enodes, bc=mesh.bc)
    s = fsb.FemShell()
    s.set_material(0.3, 1e6, 0.05)
    s.set_mesh(m["xyz"], m["etype"], m["eptr"], m["enodes"], m["bc"])
    ids = s.interface_nodes()
    assert ids.size == 43                       # fluid_solver.cpp:43-51 hard-codes 43 interface nodes
    f = np.zeros((ids.size, 2))
    f[:, 0] = 1.0 + np.sin(1 / 25.01)
    d1, _ = s.step(2, "z", f, rtol=1e-11, max_its=200000)
    F = np.zeros((mesh.n_nodes, 6))
    F[ids, 0] = f[:, 0]
    uo = fso.direct_solve(mesh, fso.assemble(mesh, F, 0.3, 1e6, 0.05))
    assert np.linalg.norm(d1 - uo[ids][:, :2]) <= 1e-7 * np.linalg.norm(uo[ids][:, :2])
    s.commit_step(2, "z")
    d2, _ = s.step(2, "z", f, rtol=1e-11, max_its=200000)   # same load after commit -> zero increment
    assert np.abs(d2).max() <= 1e-8 * np.abs(d1).max()


# ---- full-size properties (BASELINE config 2: 1000 x 1000 nodes Quad-4, 6 M DOF) ------------
def test_full_size_properties(fsb):
    n = 999
    m = fsb.meshgen("q", n, n, 0, 0, 10, 10, (1, 1, 1, 1), 300.0, 2, 1)
    s = gpu_system(fsb, m, 0.3, 1e7, 0.5, dof=0, loads=m["forces"])
    sz = s.sizes()
    assert sz["n_dofnodes"] == 1000 * 1000
    assert sz["n_blocks"] == 8988004             # 9 n^2 - 12 n + 4 ... node blocks
    rng = np.random.default_rng(3)
    x, y = rng.standard_normal(6 * sz["n_dofnodes"]), rng.standard_normal(6 * sz["n_dofnodes"])
    Ax, Ay = s.spmv(x), s.spmv(y)
    assert abs(y @ Ax - x @ Ay) <= 1e-10 * (np.linalg.norm(y) * np.linalg.norm(Ax))      # symmetry
    assert np.abs(s.spmv(2.0 * x - 3.0 * y) - (2.0 * Ax - 3.0 * Ay)).max() <= 1e-9 * np.abs(Ax).max()  # linearity
    # a rigid translation produces no force on rows away from the clamped boundary
    dn = s.dof_order()
    tz = np.zeros(6 * sz["n_dofnodes"]); tz[2::6] = 1.0
    r = s.spmv(tz).reshape(-1, 6)
    ij = np.arange(1000)
    interior = np.ones((1000, 1000), bool); interior[[0, 1, -2, -1], :] = False; interior[:, [0, 1, -2, -1]] = False
    rows = dn[np.nonzero(interior.ravel())[0]]
    assert np.abs(r[rows]).max() <= 1e-9 * np.abs(Ax).max()
    # CG decreases the energy functional phi(x) = x.Ax/2 - x.b monotonically (the residual norm is not
    # monotone on this biharmonic-like operator, so it is not the property to test)
    b = s.export_rhs()

    def phi(max_its):
        info = s.solve(rtol=1e-30, max_its=max_its, warm_start=False, allow_not_converged=True)
        assert info.iterations == max_its
        xs = np.zeros(6 * sz["n_dofnodes"])
        xs.reshape(-1, 6)[dn] = s.solution()
        return 0.5 * xs @ s.spmv(xs) - xs @ b

    p100, p300 = phi(100), phi(300)
    assert p300 < p100 < 0.0


# ---- slice pass: shells in the xy plane are assembled straight into the compacted SpMV format ----
def _planar_umbrella(mixed):
    m = dict(meshes.umbrella(mixed=mixed, n_rings=5, cone=0.0))
    xyz = np.asarray(m["xyz"], float) @ meshes.rotation(0.2, 0.4, -0.3)      # undo the generator's rigid rotation
    xyz[:, 2] = 0.0
    m["xyz"] = np.ascontiguousarray(xyz)
    return m


def _delaunay_clamped():
    m = dict(meshes.delaunay_patch(n_points=900, seed=3, quad_fraction=0.4))
    # clamp the first side of the first 40 elements' first nodes (arbitrary but deterministic Dirichlet set)
    m["bc"] = np.array([(e, 0, 1 if e % 2 else 0) for e in range(0, 40)], np.int32).reshape(-1, 3)
    F = np.zeros((m["xyz"].shape[0], 6))
    F[:, 2] = 0.01
    F[::7, 0] = 0.02
    m["forces"] = F
    return m


SLICE_CASES = {
    "c1_tri16": lambda fsb: (fsb.meshgen("t", 16, 16, 0, 0, 10, 10, (1, 1, 1, 1), 300.0, 2, 1), 0.3, 1e7, 0.5),
    "tri_ur_33x70": lambda fsb: (fsb.meshgen("t", 33, 70, -1, 0, 4, 6, (0, 21, -1, 1), 2.0, 2, 0), 0.25, 3e4, 1.0),
    "quad_41x29": lambda fsb: (fsb.meshgen("q", 41, 29, 0, 0, 10, 7, (0, 1, 20, 21), 300.0, 2, 1), 0.3, 1e7, 0.5),
    "quad_1col": lambda fsb: (fsb.meshgen("q", 1, 50, 0, 0, 1, 10, (1, 1, -1, -1), 5.0, 2, 1), 0.3, 1e7, 0.5),
    "umbrella_flat_mixed": lambda fsb: (_planar_umbrella(True), 0.3, 1e4, 0.25),
    "umbrella_flat_quads": lambda fsb: (_planar_umbrella(False), 0.3, 1e4, 0.25),
    "delaunay_mixed": lambda fsb: (_delaunay_clamped(), 0.3, 1e5, 0.1),
}


@pytest.mark.parametrize("dof", [0, 1])
@pytest.mark.parametrize("case", sorted(SLICE_CASES))
def test_slice_pass_assembles_the_compacted_format_directly(fso, fsb, case, dof):
    """fs_slice_asm.cu: no parity array is written by fs_assemble; the matrix the iteration streams is checked through
    its SpMV, its diagonal (Jacobi / block-Jacobi extraction from the sliced layout) and the solves; the parity CSR
    that fs_export_csr forms on demand still matches the oracle entry by entry"""
    m, nu, E, t = SLICE_CASES[case](fsb)
    om = as_fso_mesh(fso, m)
    ref = fso.assemble(om, m["forces"], nu, E, t, dof_mode=dof)
    s = gpu_system(fsb, m, nu, E, t, dof=dof, loads=m["forces"], asm=fsb.ASM_GATHER)
    fmt = s.spmv_format()
    assert fmt["nz_per_block"] == 14, fmt
    rng = np.random.default_rng(5)
    for _ in range(2):
        x = rng.standard_normal(6 * ref.n_dofnodes)
        y, yr = s.spmv(x), fso.spmv(ref, x)
        assert np.abs(y - yr).max() <= 1e-13 * np.abs(yr).max()
    # Dirichlet rows of the streamed matrix: unit vectors come back as integer multiples of themselves
    con = np.repeat(ref.mask[np.argsort(ref.dofnode)], 6) >> np.tile(np.arange(6), ref.n_dofnodes) & 1
    ones = np.ones(6 * ref.n_dofnodes)
    yc = s.spmv(con.astype(float))
    assert np.all(yc[con == 1] == np.round(yc[con == 1])) and np.all(yc[con == 1] >= 1)
    assert np.all((s.spmv(ones * (con == 0)))[con == 1] == 0.0)
    if np.abs(ref.rhs).max() > 0 and ref.mask.any():
        uo = fso.direct_solve(om, ref)
        for pc in (fsb.PC_JACOBI, fsb.PC_BJACOBI6):
            xo, its_o, _ = fso.pcg(ref, pc=pc, rtol=1e-10, max_its=400000)
            if its_o < 0:      # K is not positive definite for this support (drilling term, fs.cpp:1036-1051): both CGs break down
                with pytest.raises(fsb.FemShellError) as ei:
                    s.solve(rtol=1e-10, max_its=400000, pc=pc, warm_start=False)
                assert ei.value.code == fsb.FS_ERR_BREAKDOWN
                continue
            info = s.solve(rtol=1e-10, max_its=400000, pc=pc, warm_start=False)
            # the element kernels contract their sums into FMA chains, the oracle (gcc, no contraction across statements)
            # does not: values agree to 1e-12 (checked below), and on the one-column strip -- 612 unknowns, condition
            # number ~1e9 -- that last-bit difference moves the Jacobi-PCG count by a few per cent
            slack = its_o // 8 if case == "quad_1col" else its_o // 50
            assert abs(info.iterations - its_o) <= max(3, slack), (pc, info.iterations, its_o)
            assert np.linalg.norm(s.solution() - uo) <= 1e-7 * np.linalg.norm(uo)
    # on demand: the parity CSR (explicit zeros included)
    rowptr, colidx, vals = s.export_csr()
    rr, rc, rv = ref.csr()
    assert np.array_equal(rowptr, rr) and np.array_equal(colidx, rc)
    # sliver triangles of the Delaunay patch: a coupling block is a difference of terms as large as the diagonal
    # blocks, so its rounding noise scales with the element, not with the (cancelled) block (see
    # test_element_matrices_general_elements)
    tol = 1e-11 if case == "delaunay_mixed" else 1e-12
    assert block_scaled_error(vals, rv, ref.nptr) <= tol
    assert s.spmv_format()["nz_per_block"] == 14          # forming it did not change what the iteration streams
    # re-assembly with another material refreshes the compacted values (and the diagonal taken from them)
    s.set_material(0.2, 2 * E, 0.8 * t)
    s.assemble()
    ref2 = fso.assemble(om, m["forces"], 0.2, 2 * E, 0.8 * t, dof_mode=dof)
    y2, y2r = s.spmv(x), fso.spmv(ref2, x)
    assert np.abs(y2 - y2r).max() <= 1e-13 * np.abs(y2r).max()
    assert block_scaled_error(s.export_csr(with_cols=False)[2], ref2.vals, ref2.nptr) <= tol


def test_slice_pass_is_bitwise_reproducible_and_matches_the_copy_from_parity(fsb):
    m = fsb.meshgen("q", 300, 200, 0, 0, 10, 7, (1, 0, 1, 0), 300.0, 2, 1)
    x = np.random.default_rng(2).standard_normal(6 * 301 * 201)
    a = gpu_system(fsb, m, 0.3, 1e7, 0.5, asm=fsb.ASM_GATHER)
    b = gpu_system(fsb, m, 0.3, 1e7, 0.5, asm=fsb.ASM_GATHER)
    ya, yb = a.spmv(x), b.spmv(x)
    assert np.array_equal(ya, yb)
    c = gpu_system(fsb, m, 0.3, 1e7, 0.5, asm=fsb.ASM_COLORED)    # parity array first, compacted copy made from it
    assert c.spmv_format()["nz_per_block"] == 14
    yc = c.spmv(x)
    assert np.abs(ya - yc).max() <= 1e-13 * np.abs(yc).max()


def test_dmma_and_fma_contractions_agree(fsb):
    """fs_bench_contraction (north_star (a)): both variants of the plate contraction produce the same matrices"""
    r = fsb.FemShell().bench_contraction(n_elem=4096, reps=1)
    assert r["max_rel_diff"] <= 1e-13 and r["fma_ms"] > 0 and r["dmma_ms"] > 0


# ---- planar shells in general position: plane frame (slice pass + rotating SpMV) ----
@pytest.mark.parametrize("kind", ["q", "t"])
@pytest.mark.parametrize("angles", [(0.0, 0.0, np.pi / 6), (np.pi / 6, 0.0, 0.0), (0.3, -0.5, 0.7)], ids=["about_z", "about_x_30deg", "general"])
def test_rotated_plate_uses_the_compacted_format_in_its_plane_frame(fso, fsb, kind, angles):
    """a plate rotated out of the coordinate planes streams 14 of 36 entries per block too (the
    slice pass assembles Q~ K Q~^T in the plane frame, the SpMV rotates x blocks in and y blocks out); products,
    diagonals and solves agree with the oracle, which knows nothing of the frame"""
    m = fsb.meshgen(kind, 37, 21, 0, 0, 10, 6, (1, 0, 1, 20), 300.0, 2, 1)
    R = meshes.rotation(*angles)
    m = _rotate_mesh(m, R)
    om = as_fso_mesh(fso, m)
    ref = fso.assemble(om, m["forces"], 0.3, 1e7, 0.5)
    s = gpu_system(fsb, m, 0.3, 1e7, 0.5, loads=m["forces"], asm=fsb.ASM_GATHER)
    assert s.spmv_format()["nz_per_block"] == 14
    rng = np.random.default_rng(17)
    for _ in range(2):
        x = rng.standard_normal(6 * ref.n_dofnodes)
        y, yr = s.spmv(x), fso.spmv(ref, x)
        assert np.abs(y - yr).max() <= 1e-12 * np.abs(yr).max()
    uo = fso.direct_solve(om, ref)
    for pc in (fsb.PC_JACOBI, fsb.PC_BJACOBI6):
        xo, its_o, _ = fso.pcg(ref, pc=pc, rtol=1e-10, max_its=400000)
        info = s.solve(rtol=1e-10, max_its=400000, pc=pc, warm_start=False)
        assert abs(info.iterations - its_o) <= max(3, its_o // 50), (pc, info.iterations, its_o)
        assert np.linalg.norm(s.solution() - uo) <= 1e-7 * np.linalg.norm(uo)
    info = s.solve(rtol=1e-10, max_its=3000, pc=fsb.PC_MLRBM, warm_start=False)
    assert np.linalg.norm(s.solution() - uo) <= 1e-7 * np.linalg.norm(uo)
    # the parity CSR is formed from the ORIGINAL coordinates on demand
    rowptr, colidx, vals = s.export_csr()
    rr, rc, rv = ref.csr()
    assert np.array_equal(rowptr, rr) and np.array_equal(colidx, rc)
    assert block_scaled_error(vals, rv, ref.nptr) <= 1e-12
    # and the parity-format kernel gives the same product
    s.set_spmv_format(fsb.SPMV_FULL)
    assert s.spmv_format()["nz_per_block"] == 36
    assert np.abs(s.spmv(x) - y).max() <= 1e-12 * np.abs(yr).max()

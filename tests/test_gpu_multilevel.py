"""GPU tests of the multilevel preconditioner FS_PC_MLRBM (no counterpart in the reference; it must leave the
converged displacements of the reference path untouched and only change the iteration count)."""
import numpy as np
import pytest

import meshes
from ml_mirror import Mirror, element_extent
from test_gpu_parity import as_fso_mesh, gpu_system

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fsb():
    import fem_shell_b200 as fsb
    return fsb


CASES = {
    "quad40": lambda fsb: (fsb.meshgen("q", 40, 33, 0, 0, 10, 8.25, (1, 1, 1, 1), 300.0, 2, 1), 0.3, 1e7, 0.5),
    "tri31": lambda fsb: (fsb.meshgen("t", 31, 31, 0, 0, 10, 10, (1, 0, 1, 0), 300.0, 2, 1), 0.3, 1e7, 0.25),
    "tri_xz": lambda fsb: (fsb.meshgen("t", 20, 26, -1, 0, 4, 6.5, (0, 0, -1, 1), 2.0, 1, 0, "y"), 0.25, 3e4, 0.3),
    "folded": lambda fsb: (meshes.folded_cantilever(nx=36, ny=12), 0.3, 1e4, 0.25),
}


def dof_ordered(ref, m):
    xyz = np.zeros((ref.n_dofnodes, 3))
    xyz[ref.dofnode] = np.asarray(m["xyz"], float)
    mask = np.zeros(ref.n_dofnodes, np.uint8)
    mask[ref.dofnode] = ref.mask
    return xyz, mask


@pytest.mark.parametrize("gamma", [1, 2, 21])      # 21: two visits of the second lattice per visit of the first, one below
@pytest.mark.parametrize("case", sorted(CASES))
def test_cycle_matches_explicit_galerkin_mirror(fso, fsb, case, gamma):
    """the probed stencils, transfer kernels, smoothers and the dense coarsest solve reproduce the cycle built
    from explicit sparse Galerkin products with the same smoother weights"""
    m, nu, E, t = CASES[case](fsb)
    om = as_fso_mesh(fso, m)
    ref = fso.assemble(om, m["forces"], nu, E, t)
    s = gpu_system(fsb, m, nu, E, t, loads=m["forces"])
    s.set_ml_options(dense_points=24, gamma=gamma)      # small dense level -> several lattice levels
    rng = np.random.default_rng(7)
    r = rng.standard_normal(6 * ref.n_dofnodes)
    z = s.apply_mlrbm(r)
    info = s.ml_info()
    assert info["levels"] >= 2 and info["setup_ms"] > 0
    xyz, mask = dof_ordered(ref, m)
    h = element_extent(np.asarray(m["xyz"], float), m["eptr"], m["enodes"])
    M = Mirror(ref.scipy().tocsr(), xyz, mask, h, info["cells"], info["lambda"], gamma=gamma)
    # the power-iteration estimates (x 1.1) bound the spectra they stand for from above, loosely
    for l in range(info["levels"]):
        lam = M.lambda_max(l)
        assert 0.9 * lam <= info["lambda"][l] <= 1.25 * lam, (l, lam, info["lambda"][l])
    zr = M(r)
    assert np.linalg.norm(z - zr) <= 1e-9 * np.linalg.norm(zr), np.linalg.norm(z - zr) / np.linalg.norm(zr)
    # symmetric and positive definite, as CG needs
    r2 = rng.standard_normal(r.size)
    z2 = s.apply_mlrbm(r2)
    assert abs(r2 @ z - r @ z2) <= 1e-9 * abs(r @ z2) + 1e-12 * np.linalg.norm(r) * np.linalg.norm(z2)
    assert r @ z > 0 and r2 @ z2 > 0


@pytest.mark.parametrize("case", sorted(CASES))
def test_pcg_with_multilevel_reaches_the_reference_displacements(fso, fsb, case):
    m, nu, E, t = CASES[case](fsb)
    om = as_fso_mesh(fso, m)
    ref = fso.assemble(om, m["forces"], nu, E, t)
    uo = fso.direct_solve(om, ref)
    s = gpu_system(fsb, m, nu, E, t, loads=m["forces"])
    jac = s.solve(rtol=1e-11, max_its=400000, pc=fsb.PC_JACOBI, warm_start=False)
    ml = s.solve(rtol=1e-11, max_its=2000, pc=fsb.PC_MLRBM, warm_start=False)
    assert ml.status == 0 and ml.rel_residual <= 1e-11
    u = s.solution()
    assert np.linalg.norm(u - uo) <= 1e-8 * np.linalg.norm(uo)
    assert ml.iterations * 4 < jac.iterations, (ml.iterations, jac.iterations)
    # values pass again (same values): the set-up is redone, the captured iteration is not reused stale
    s.assemble()
    again = s.solve(rtol=1e-11, max_its=2000, pc=fsb.PC_MLRBM, warm_start=False)
    assert again.iterations == ml.iterations
    # and the other preconditioners still work on the same context afterwards
    j2 = s.solve(rtol=1e-11, max_its=400000, pc=fsb.PC_JACOBI, warm_start=False)
    assert j2.iterations == jac.iterations


def test_multilevel_iterations_grow_slowly(fsb):
    its = []
    for n in (32, 64, 128):
        m = fsb.meshgen("q", n, n, 0, 0, 10, 10, (1, 1, 1, 1), 300.0, 2, 1)
        s = gpu_system(fsb, m, 0.3, 1e7, 0.5, loads=m["forces"])
        its.append(s.solve(rtol=1e-8, max_its=2000, pc=fsb.PC_MLRBM, warm_start=False).iterations)
    assert its[0] <= 60 and its[2] <= 110 and its[2] <= 2.2 * its[0], its


def test_stage_profile_of_the_multilevel_iteration(fsb, monkeypatch):
    """FS_ML_PROFILE=1: the iterations run eagerly between events; same iteration count as the captured graph, and the
    stage times add up to (about) the solve time"""
    m = fsb.meshgen("q", 90, 70, 0, 0, 10, 8, (1, 1, 1, 1), 300.0, 2, 1)
    s = gpu_system(fsb, m, 0.3, 1e7, 0.5, loads=m["forces"])
    s.set_ml_options(dense_points=24)
    ref = s.solve(rtol=1e-9, max_its=2000, pc=fsb.PC_MLRBM, warm_start=False)
    u = s.solution()
    monkeypatch.setenv("FS_ML_PROFILE", "1")
    s.ml_profile(reset=True)
    info = s.solve(rtol=1e-9, max_its=2000, pc=fsb.PC_MLRBM, warm_start=False)
    p = s.ml_profile()
    assert info.iterations == ref.iterations == p["iterations"]
    assert np.array_equal(s.solution(), u)
    stages = [p[k] for k in ("halo_spmv_update", "presmooth_restrict", "lattice_cycle", "prolong", "postsmooth_spmv_rz", "allreduce_direction")]
    assert all(v > 0 for v in stages)
    assert 0.5 * info.solve_ms <= sum(stages) * info.iterations <= 1.05 * info.solve_ms
    assert p["lattice_level2_visit1"] + p["lattice_level2_visit2"] <= p["lattice_cycle"]


@pytest.mark.parametrize("case", ["quad40", "tri_xz", "folded"])
def test_compacted_lattice_stencils_change_no_bit(fsb, case, monkeypatch):
    """shells in a coordinate plane iterate on 18 of the 36 entries of every lattice block (k_lat_stencil_c); the skipped
    entries are exact zeros and the sums keep their order, so z = M^-1 r is bit-identical to the full-block kernels.  The
    folded shell is not planar: no level is compacted"""
    m, nu, E, t = CASES[case](fsb)
    monkeypatch.setenv("FS_ML_COMPACT_MIN_CELLS", "1")     # by default only lattices that stream from HBM are compacted
    s = gpu_system(fsb, m, nu, E, t, loads=m["forces"])
    s.set_ml_options(dense_points=24)
    r = np.random.default_rng(11).standard_normal(6 * s.sizes()["n_dofnodes"])
    z1 = s.apply_mlrbm(r)
    info = s.ml_info()
    assert info["compact_levels"] == (0 if case == "folded" else info["levels"] - 1), info
    monkeypatch.setenv("FS_ML_COMPACT", "0")
    s.assemble()                                        # new values -> the set-up runs again, now without compaction
    z0 = s.apply_mlrbm(r)
    assert s.ml_info()["compact_levels"] == 0
    assert np.array_equal(z0, z1)


def test_multilevel_rejects_preconditioned_norm(fsb):
    m = fsb.meshgen("q", 8, 8, 0, 0, 10, 10, (1, 1, 1, 1), 300.0, 2, 1)
    s = gpu_system(fsb, m, 0.3, 1e7, 0.5, loads=m["forces"])
    with pytest.raises(fsb.FemShellError):
        s.solve(rtol=1e-8, pc=fsb.PC_MLRBM, norm_type=fsb.NORM_PRECONDITIONED)
    info = s.solve(rtol=1e-8, max_its=500, pc=fsb.PC_MLRBM)     # tiny mesh: the first lattice is already dense
    assert info.status == 0

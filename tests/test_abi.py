"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/femshell_b200.h declares, its host-side functions (no GPU needed) agree with the oracle's
restatements bit for bit, and the product refuses to run without a CUDA device."""
import os
import re

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

import fem_shell_b200 as fsb


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "femshell_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fs_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = fsb.load_library()
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), "missing export " + n
    assert sorted(fsb.EXPORTED_SYMBOLS) == names


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(fsb.FemShellError):
        fsb.FemShell(device=0)


def test_product_does_not_import_oracle():
    for root, _, files in os.walk(os.path.join(ROOT, "fem_shell_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                txt = open(os.path.join(root, f)).read()
                assert "fso" not in txt.replace("fso_", "") or f == "__none__", f
                assert "import oracle" not in txt and "from oracle" not in txt and "libfs_oracle" not in txt, f


import importlib.util
_spec = importlib.util.spec_from_file_location("make_fixtures", os.path.join(GOLDEN, "make_fixtures.py"))
_mf = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_mf)


@pytest.mark.parametrize("name", sorted(_mf.MESHGEN_CASES))
def test_cxx_meshgen_matches_reference_generator(name):
    z = np.load(os.path.join(GOLDEN, "meshgen_ref.npz"))
    m = fsb.meshgen(*_mf.MESHGEN_CASES[name])
    assert np.array_equal(m["xyz"], z[name + "/xyz"])
    assert np.array_equal(m["etype"], z[name + "/etype"])
    assert np.array_equal(m["enodes"], z[name + "/enodes"])
    assert np.array_equal(m["bc"].reshape(-1, 3), z[name + "/bc"].reshape(-1, 3))
    assert np.array_equal(m["forces"], z[name + "/forces"])


def test_cxx_meshgen_matches_oracle_restatement(fso):
    for a in (("q", 13, 7, 0.0, 0.0, 10.0, 3.0, (1, 0, 20, 21), 117.0, 2, 1, "z"),
              ("t", 999, 3, 0.0, 0.0, 10.0, 10.0, (1, 1, 1, 1), 300.0, 2, 1, "z"),
              ("t", 11, 17, -2.0, 1.0, 5.0, 9.0, (-1, 2, 0, 1), 0.125, 1, 0, "x")):
        m = fsb.meshgen(*a)
        om, F = fso.meshgen(*a)
        assert np.array_equal(m["xyz"], om.xyz) and np.array_equal(m["enodes"], om.enodes)
        assert np.array_equal(m["bc"], om.bc) and np.array_equal(m["forces"], F)


def test_cxx_xda_and_force_readers(fso, ref_meshes, tmp_path):
    from conftest import load_ref_mesh
    for name in ("test_A_uv_t", "test_E_uvw_t", "test_D_w_q_uni16"):
        mesh, F = load_ref_mesh(fso, ref_meshes, name)
        p = str(tmp_path / (name + ".xda"))
        fsb.write_xda(p, mesh.xyz, mesh.etype, mesh.eptr, mesh.enodes, mesh.bc)
        back = fsb.read_xda(p)
        assert np.array_equal(back["xyz"], mesh.xyz) and np.array_equal(back["enodes"], mesh.enodes)
        assert np.array_equal(back["etype"], mesh.etype) and np.array_equal(back["bc"], mesh.bc)
        ob = fso.read_xda(p)                       # the oracle's reader sees the same file
        assert np.array_equal(ob.xyz, mesh.xyz) and np.array_equal(ob.bc, mesh.bc)
    # load file with n-1 rows under a header of n (main_all.cpp:352,377): last node stays zero
    pf = str(tmp_path / "m_f")
    with open(pf, "w") as f:
        f.write("4\n2.5\n0 0 1 0 0 0\n0 0 1 0 0 0\n1 0 0 0 0 0.5\n")
    F = fsb.read_forces(pf, 4)
    assert np.array_equal(F, fso.read_forces(pf, 4))
    assert F[3].tolist() == [0] * 6 and F[2].tolist() == [2.5, 0, 0, 0, 0, 1.25]
    with pytest.raises(fsb.FemShellError):
        fsb.read_xda(str(tmp_path / "missing.xda"))

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


def test_text_readers_tolerate_crlf_comments_and_short_files(tmp_path):
    """the whole-file cursor behind fs_read_xda / fs_read_forces: CRLF line ends, comment-only and blank lines, no trailing
    newline; a load file that ends early leaves the remaining entries zero like the reference's failed stream extraction
    (fs.cpp:59-66)"""
    xda = tmp_path / "edge.xda"
    xda.write_bytes(b"libMesh-0.7.0+\r\n2   # elems\r\n\r\n5 # nodes\r\n.  # bc file\r\nn/a\r\nn/a # p\r\nn/a\r\n2 # level 0\r\n"
                    b"3 0 1 2\r\n# a comment line\r\n5 1 3 4 2 # quad\r\n0 0 0\r\n1 0 0\r\n0 1 0\r\n1 1 0.5\r\n2e0 1 -.25\r\n1 # nbc\r\n0 2 1")
    r = fsb.read_xda(str(xda))
    assert r["etype"].tolist() == [3, 5] and r["eptr"].tolist() == [0, 3, 7] and r["enodes"].tolist() == [0, 1, 2, 1, 3, 4, 2]
    assert r["bc"].tolist() == [[0, 2, 1]]
    assert np.array_equal(r["xyz"], np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0.5], [2, 1, -0.25]], float))
    ff = tmp_path / "edge_f"
    ff.write_text("5\n2.0\n1 2 3 4 5 6\n7 8 9 10 11 12\n1 1\n")
    F = fsb.read_forces(str(ff), 5)
    assert np.array_equal(F, np.array([[2, 4, 6, 8, 10, 12], [14, 16, 18, 20, 22, 24], [2, 2, 0, 0, 0, 0], [0] * 6, [0] * 6], float))
    with pytest.raises(fsb.FemShellError):
        fsb.read_xda(str(tmp_path / "missing.xda"))
    bad = tmp_path / "bad.xda"
    bad.write_text("libMesh-0.7.0+\n1\n3\n.\nn/a\nn/a\nn/a\n1\n10 0 1 2 3 4 5 6 7\n")   # HEX8: not an element fem-shell handles
    with pytest.raises(fsb.FemShellError):
        fsb.read_xda(str(bad))


def test_xda_writer_round_trip_and_reference_text(tmp_path):
    m = fsb.meshgen("t", 7, 5, -1.5, 0.25, 3.0, 2.0, (0, 1, 20, 21), 3.0, 2, 1)
    p = str(tmp_path / "rt.xda")
    fsb.write_xda(p, m["xyz"], m["etype"], m["eptr"], m["enodes"], m["bc"])
    r = fsb.read_xda(p)
    for k in ("xyz", "etype", "eptr", "enodes"):
        assert np.array_equal(r[k], m[k]), k
    assert np.array_equal(r["bc"].reshape(-1, 3), m["bc"].reshape(-1, 3))
    lines = open(p).read().splitlines()
    assert lines[0] == "libMesh-0.7.0+" and lines[8] == "3 " + " ".join(str(v) for v in m["enodes"][:3])
    assert lines[8 + m["etype"].size] == "%g %g %g" % tuple(m["xyz"][0])

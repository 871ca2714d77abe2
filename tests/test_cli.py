"""Command-line twins of the reference's executables (fem_shell_b200/bin): meshGen byte-for-byte against
the reference generator (CPU), fem-shell / fem-shell-coupled end to end on the GPU."""
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_ref_mesh

import fem_shell_b200 as fsb

BIN = os.path.join(ROOT, "fem_shell_b200", "bin")
REF_MESHGEN = os.path.join(ROOT, "oracle", "_ref", "meshgen")

ARGSETS = [
    "q 16 16 0 0 10 10 0,0,0,0 300 2 1 z",
    "t 7 9 -1 0.5 2 3 1,-1,2,0 2.5 2 0 y",
    "t 12 5 0 0 10 2 0,1,20,21 30000 1 1 z",
    "q 33 3 0 0 1 1 -1,1,-1,-1 1 0 1 x",
    "q 999 2 0 0 10 10 1,1,1,1 300 2 1 z",
]


@pytest.mark.parametrize("argline", ARGSETS)
def test_meshgen_cli_is_byte_identical_to_reference(argline, tmp_path):
    if not os.path.exists(REF_MESHGEN):
        pytest.skip("oracle/_ref/meshgen not built (reference tree absent)")
    a, b = str(tmp_path / "ours"), str(tmp_path / "ref")
    assert subprocess.run([os.path.join(BIN, "meshGen")] + argline.split() + [a]).returncode == 0
    assert subprocess.run([REF_MESHGEN] + argline.split() + [b]).returncode == 0
    assert open(a + ".xda", "rb").read() == open(b + ".xda", "rb").read()
    if int(argline.split()[9]) > 0:
        assert open(a + "_f", "rb").read() == open(b + "_f", "rb").read()
    else:
        assert not os.path.exists(a + "_f")


def test_cli_usage_errors():
    assert subprocess.run([os.path.join(BIN, "meshGen"), "q", "1"], capture_output=True).returncode != 0
    r = subprocess.run([os.path.join(BIN, "fem-shell"), "-nu", "0.3"], capture_output=True, text=True)
    assert r.returncode != 0 and "FAILED" in r.stdout          # fs.cpp:21-25
    r = subprocess.run([os.path.join(BIN, "fem-shell"), "-nu", "0.3", "-e", "1", "-t", "1", "-out", "x"], capture_output=True, text=True)
    assert r.returncode != 0 and "Mesh file not specified" in r.stderr


def _solution_rows(stdout):
    rows = re.findall(r"u= (\S+), v= (\S+), w= (\S+), tx= (\S+), ty= (\S+), tz= (\S+)\]", stdout)
    return np.array(rows, float)


@pytest.mark.gpu
def test_fem_shell_cli_reproduces_thesis_values(fso, ref_meshes, tmp_path):
    """run_examples.sh Test D and Test A through the stand-alone binary, values read back from its stdout"""
    for name, flags, checks in (
        ("test_D_w_q_uni16", ["-nu", "0.3", "-e", "1e7", "-t", "0.5"], [(144, 2, 0.106454)]),
        ("test_A_uv_t", ["-nu", "0.25", "-e", "30000", "-t", "1.0"], [(22, 0, -0.0255988), (26, 1, 0.194407)]),
    ):
        mesh, F = load_ref_mesh(fso, ref_meshes, name)
        base = str(tmp_path / name)
        fsb.write_xda(base + ".xda", mesh.xyz, mesh.etype, mesh.eptr, mesh.enodes, mesh.bc)
        with open(base + "_f", "w") as f:       # fs.cpp:52-66 format
            f.write("%d\n1.0\n" % mesh.n_nodes)
            for row in F:
                f.write(" ".join(repr(float(v)) for v in row) + "\n")
        for pc in ("pbjacobi", "mg"):   # the reference's documented block Jacobi, and the multilevel cycle
            r = subprocess.run([os.path.join(BIN, "fem-shell")] + flags + ["-mesh", base + ".xda", "-out", base, "-pc_type", pc,
                                "-ksp_max_it", "100000"], capture_output=True, text=True, timeout=300)
            assert r.returncode == 0, r.stdout[-2000:] + r.stderr
            assert "Read command-line arguments.......OK" in r.stdout and "All done :)" in r.stdout
            u = _solution_rows(r.stdout)
            assert u.shape == (mesh.n_nodes, 6)
            for node, var, gold in checks:
                assert float("%.6g" % u[node, var]) == pytest.approx(gold, rel=2e-6)
        assert os.path.exists(base + ".vtk")
        vtk = open(base + ".vtk").read()
        assert "POINTS %d double" % mesh.n_nodes in vtk and "SCALARS tz double 1" in vtk
        assert "CELL_DATA %d" % mesh.n_elem in vtk and "SCALARS M_xy double 1" in vtk


@pytest.mark.gpu
def test_coupled_cli_runs_the_tower(fso, ref_meshes, tmp_path):
    mesh, _ = load_ref_mesh(fso, ref_meshes, "bending_tower_tri_test")
    base = str(tmp_path / "tower")
    fsb.write_xda(base + ".xda", mesh.xyz, mesh.etype, mesh.eptr, mesh.enodes, mesh.bc)
    r = subprocess.run([os.path.join(BIN, "fem-shell-coupled"), "-nu", "0.3", "-e", "1e6", "-t", "0.05", "-mesh", base + ".xda",
                        "-config", "precice_config.xml", "-dt", "0.01", "-axis", "z", "-steps", "4", "-subiters", "2",
                        "-ksp_max_it", "200000", "-ksp_rtol", "1e-10"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr
    assert "coupling interface nodes = 43" in r.stdout
    assert r.stdout.count("Advancing in time, finished timestep") == 4 and r.stdout.count("Iterate") == 4
    inc = [float(x) for x in re.findall(r"max \|increment\| (\S+)\)", r.stdout)]
    # linear problem, load amplitude 1+sin(t/25.01): step 0 carries the whole displacement, step 1 adds sin(1/25.01) of it
    assert inc[1] > 0 and abs(inc[1] / inc[0] - np.sin(1 / 25.01)) < 1e-4

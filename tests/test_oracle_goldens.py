"""Pins the CPU oracle (oracle/fs_oracle.c) to every golden the reference publishes for the
assembly->solve path: the thesis tables of doc/validation.tex (Tests A-G), through the shipped
example inputs and through meshGen-semantics regenerated inputs.  Tolerance = the printed
precision (6 significant digits => 1e-5 relative)."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_ref_mesh

G = json.load(open(os.path.join(GOLDEN, "thesis_goldens.json")))


def agree6(value, gold):
    """value printed with 6 significant digits (std::cout default, fs.cpp:166) equals gold"""
    import math
    ulp6 = 10.0 ** (math.floor(math.log10(abs(gold))) - 5)
    return abs(value - gold) <= 0.5001 * ulp6


@pytest.mark.parametrize("case", G["shipped"], ids=[c["test"] for c in G["shipped"]])
def test_shipped_examples(fso, ref_meshes, case):
    mesh, F = load_ref_mesh(fso, ref_meshes, case["mesh"])
    s = fso.assemble(mesh, F, case["nu"], case["E"], case["t"])
    u = fso.direct_solve(mesh, s)
    for node, var, gold in case["checks"]:
        assert agree6(u[node, var], gold), (case["test"], node, var, u[node, var], gold, case["cite"])


def _rows():
    out = []
    for name, blk in G["meshgen"].items():
        if name.startswith("_"):
            continue
        for r in blk["rows"]:
            out.append(pytest.param(blk, r, id="%s-N%d-l%d-ul%d" % (name, r[0], r[1], r[3])))
    return out


@pytest.mark.parametrize("blk,row", _rows())
def test_meshgen_examples(fso, blk, row):
    N, loading, factor, ul, gold = row
    mesh, F = fso.meshgen(blk["kind"], N, N, 0.0, 0.0, blk["Lx"], blk["Ly"], (blk["bc"],) * 4, factor, loading, ul)
    s = fso.assemble(mesh, F, blk["nu"], blk["E"], blk["t"])
    u = fso.direct_solve(mesh, s)
    assert agree6(u[mesh.n_nodes // 2, 2], gold), (u[mesh.n_nodes // 2, 2], gold, blk["cite"])


def test_quirks_silent_on_fixtures(fso, ref_meshes):
    """none of the shipped inputs exercises fs.cpp:586 / the det() side effect"""
    for name, nu, E, t in (("test_C_w_tA16", 0.3, 10.92, 1.0), ("test_D_w_q_uni16", 0.3, 1e7, 0.5)):
        mesh, F = load_ref_mesh(fso, ref_meshes, name)
        a = fso.assemble(mesh, F, nu, E, t, quirks=fso.QUIRKS_REFERENCE)
        b = fso.assemble(mesh, F, nu, E, t, quirks=0)
        assert np.array_equal(a.vals, b.vals)


def test_quirks_bite_on_general_elements(fso):
    tri = np.array([[0, 0, 0], [2.0, 0.1, 0.3], [0.7, 1.5, -0.2]])
    quad = np.array([[0, 0, 0], [2.0, 0.2, 0], [2.6, 1.7, 0], [-0.3, 1.1, 0]])
    for et, X in ((fso.TRI3, tri), (fso.QUAD4, quad)):
        a = fso.element_stiffness(et, X, 0.3, 1e7, 0.5, quirks=fso.QUIRKS_REFERENCE)
        b = fso.element_stiffness(et, X, 0.3, 1e7, 0.5, quirks=0)
        assert np.abs(a - b).max() > 1e-4 * np.abs(a).max()


def test_matrix_structure(fso, ref_meshes):
    mesh, F = load_ref_mesh(fso, ref_meshes, "test_E_uvw_t")
    s = fso.assemble(mesh, F, 0.3, 1e4, 0.25)
    A = s.scipy()
    assert abs(A - A.T).max() <= 1e-12 * abs(A).max()
    # constrained rows: unit-per-element diagonal, nothing else, zero rhs (libMesh constraint, fs.cpp:1227)
    dn = s.dofnode
    for n in np.nonzero(s.mask)[0]:
        for v in range(6):
            if s.mask[n] >> v & 1:
                r = 6 * dn[n] + v
                row = A.getrow(r).toarray().ravel()
                nel = sum(1 for e in range(mesh.n_elem) if n in mesh.enodes[mesh.eptr[e]:mesh.eptr[e + 1]])
                assert row[r] == nel and np.count_nonzero(row) == 1 and s.rhs[r] == 0.0


def test_pcg_matches_direct(fso, ref_meshes):
    mesh, F = load_ref_mesh(fso, ref_meshes, "test_D_w_q_uni16")
    s = fso.assemble(mesh, F, 0.3, 1e7, 0.5)
    u = fso.direct_solve(mesh, s)
    for pc in (fso.PC_JACOBI, fso.PC_BJACOBI6):
        x, its, rel = fso.pcg(s, pc=pc, rtol=1e-10)
        assert 0 < its < 400 and rel <= 1e-10
        uu = fso.gather_solution(mesh, s, x)
        assert np.linalg.norm(uu - u) <= 1e-8 * np.linalg.norm(u)


def test_dof_orders(fso, ref_meshes):
    mesh, F = load_ref_mesh(fso, ref_meshes, "test_B_uv_q")
    d0, n0 = fso.dof_order(mesh, fso.DOF_FIRST_ENCOUNTER)
    d1, n1 = fso.dof_order(mesh, fso.DOF_NODE_ID)
    assert n0 == n1 == mesh.n_nodes
    assert list(d1) == list(range(mesh.n_nodes))
    # first element 0 1 10 9 -> those nodes get 0,1,2,3
    first = mesh.enodes[:4]
    assert [d0[i] for i in first] == [0, 1, 2, 3]
    ua = fso.direct_solve(mesh, fso.assemble(mesh, F, 0.25, 3e4, 1.0, dof_mode=0))
    ub = fso.direct_solve(mesh, fso.assemble(mesh, F, 0.25, 3e4, 1.0, dof_mode=1))
    assert np.allclose(ua, ub, rtol=1e-9, atol=1e-14)

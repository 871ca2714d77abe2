"""The two other mesh formats fem-shell accepts by extension (fs.cpp:37,45-48: *.xdr, *.msh).  The reference ships no
fixture of either, so these tests pin the readers to the format descriptions: a hand-written Gmsh file whose content is
known node by node, a Gmsh rendering of meshGen plates compared with the XDA path, and the XDR round trip (byte layout
checked against struct.pack)."""
import struct

import numpy as np
import pytest

import fem_shell_b200 as fsb


def write_msh(path, m, id_offset=1, shuffle_lines=False):
    """Gmsh 2.2 ASCII rendering of a mesh dict: boundary records become 2-node line elements with the id as physical tag"""
    xyz, etype, eptr, enodes, bc = m["xyz"], m["etype"], m["eptr"], m["enodes"], m["bc"]
    out = ["$MeshFormat", "2.2 0 8", "$EndMeshFormat", "$Nodes", str(len(xyz))]
    gid = lambda n: 3 * int(n) + id_offset            # ids with gaps, as Gmsh may produce
    for i, x in enumerate(xyz):
        out.append("%d %.17g %.17g %.17g" % (gid(i), x[0], x[1], x[2]))
    out += ["$EndNodes", "$Elements", str(len(etype) + len(bc))]
    k = 1
    lines = []
    for e, s, bid in bc:
        nen = eptr[e + 1] - eptr[e]
        n0, n1 = enodes[eptr[e] + s], enodes[eptr[e] + (s + 1) % nen]
        if shuffle_lines:
            n0, n1 = n1, n0
        lines.append((bid, n0, n1))
    for bid, n0, n1 in lines:
        out.append("%d 1 2 %d %d %d %d" % (k, bid, bid, gid(n0), gid(n1)))
        k += 1
    for e in range(len(etype)):
        nd = enodes[eptr[e]:eptr[e + 1]]
        out.append("%d %d 2 7 7 %s" % (k, 2 if len(nd) == 3 else 3, " ".join(str(gid(n)) for n in nd)))
        k += 1
    out += ["$EndElements", ""]
    open(path, "w").write("\n".join(out))


def test_gmsh_hand_written(tmp_path):
    p = tmp_path / "two.msh"
    p.write_text("""$MeshFormat
2.2 0 8
$EndMeshFormat
$PhysicalNames
2
1 20 "clamped"
2 7 "shell"
$EndPhysicalNames
$Nodes
6
10 0 0 0
11 1 0 0
12 2 0 0.5
20 0 1 0
21 1 1 0
22 2 1 0.5
$EndNodes
$Elements
5
1 15 2 99 99 10
2 1 2 20 1 20 10
3 1 2 1 1 11 12
4 3 2 7 1 10 11 21 20
5 2 2 7 1 11 12 22
$EndElements
""")
    m = fsb.read_mesh(str(p))
    assert m["xyz"].shape == (6, 3) and np.allclose(m["xyz"][2], [2, 0, 0.5]) and np.allclose(m["xyz"][4], [1, 1, 0])
    assert m["etype"].tolist() == [5, 3] and m["eptr"].tolist() == [0, 4, 7]
    assert m["enodes"].tolist() == [0, 1, 4, 3, 1, 2, 5]
    # line 20-10 = nodes (3, 0) = side 3 of the quad; line 11-12 = nodes (1, 2) = side 0 of the triangle
    assert m["bc"].tolist() == [[0, 3, 20], [1, 0, 1]]


@pytest.mark.parametrize("kind", ["q", "t"])
def test_gmsh_matches_xda_path(tmp_path, kind):
    m = fsb.meshgen(kind, 7, 5, 0.0, 0.0, 3.0, 2.0, (1, 0, 20, -1), 10.0, 2, 1)
    p = str(tmp_path / "plate.msh")
    write_msh(p, m, id_offset=5, shuffle_lines=True)
    r = fsb.read_mesh(p)
    for k in ("xyz", "etype", "eptr", "enodes"):
        assert np.array_equal(r[k], m[k]), k
    assert sorted(map(tuple, r["bc"].tolist())) == sorted(map(tuple, m["bc"].tolist()))


def test_gmsh_refusals(tmp_path):
    p = tmp_path / "bad.msh"
    p.write_text("$MeshFormat\n4.1 0 8\n$EndMeshFormat\n")
    with pytest.raises(fsb.FemShellError):
        fsb.read_mesh(str(p))
    p.write_text("$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n1\n1 0 0 0\n$EndNodes\n$Elements\n1\n1 4 2 0 0 1 1 1 1\n$EndElements\n")
    with pytest.raises(fsb.FemShellError):   # a tetrahedron: not a shell element (fs.cpp:315,342)
        fsb.read_mesh(str(p))
    with pytest.raises(fsb.FemShellError):
        fsb.read_mesh(str(tmp_path / "missing.msh"))


def test_xdr_round_trip_and_layout(tmp_path):
    m = fsb.meshgen("t", 4, 3, -1.0, 0.0, 1.0, 2.5, (1, 1, 0, 2), 10.0, 2, 0)
    p = str(tmp_path / "m.xdr")
    fsb.write_xdr(p, m["xyz"], m["etype"], m["eptr"], m["enodes"], m["bc"])
    r = fsb.read_mesh(p)
    for k in ("xyz", "etype", "eptr", "enodes", "bc"):
        assert np.array_equal(r[k], m[k]), k
    raw = open(p, "rb").read()
    ne, nn = m["etype"].size, m["xyz"].shape[0]
    head = struct.pack(">I", 14) + b"libMesh-0.7.0+\0\0" + struct.pack(">II", ne, nn)
    head += struct.pack(">I", 1) + b".\0\0\0" + (struct.pack(">I", 3) + b"n/a\0") * 3 + struct.pack(">I", ne)
    assert raw[:len(head)] == head
    first = struct.pack(">IIII", 3, *m["enodes"][:3])
    assert raw[len(head):len(head) + 16] == first
    off = len(head) + 4 * (ne + m["enodes"].size)
    assert raw[off:off + 24] == struct.pack(">ddd", *m["xyz"][0])
    assert len(raw) == off + 24 * nn + 4 + 12 * m["bc"].shape[0]
    # the same mesh through the text twin
    q = str(tmp_path / "m.xda")
    fsb.write_xda(q, m["xyz"], m["etype"], m["eptr"], m["enodes"], m["bc"])
    t = fsb.read_mesh(q)
    assert np.array_equal(t["enodes"], r["enodes"]) and np.array_equal(t["bc"], r["bc"]) and np.allclose(t["xyz"], r["xyz"], rtol=1e-5)
    open(p, "wb").write(raw[:off + 7])        # truncated inside the coordinates
    with pytest.raises(fsb.FemShellError):
        fsb.read_mesh(p)

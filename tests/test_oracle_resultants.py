"""Pins the oracle's stress-resultant recovery (oracle/fs_oracle.c fso_recover_resultants; thesis
doc/shellelements.tex:524 and :1394-1403 -- the reference ships the formulas but no code, SURVEY.md
section 8 f4) against closed-form fields both plate elements reproduce exactly: a linear in-plane
displacement (constant membrane strain) and a quadratic deflection with Kirchhoff rotations
(constant curvature).  Tolerance 1e-9 relative: the fields are inside the elements' spaces."""
import numpy as np
import pytest

NU, E, T = 0.3, 1.0e7, 0.5
KX, KY, KXY = 0.01, -0.02, 0.005
EXX, EYY, GXY = 1.0e-3, 2.5e-3, 1.5e-3


def material():
    fm = E / (1 - NU * NU)
    Dm = fm * np.array([[1, NU, 0], [NU, 1, 0], [0, 0, (1 - NU) / 2]])
    return Dm, Dm * T ** 3 / 12


def patch_field(xyz):
    """sols[node, var] of the closed-form field on a plate in the xy plane; the elements' rotation unknowns
    are theta_x = dw/dy, theta_y = -dw/dx (rigid-body tilts give zero curvature only with this pairing)"""
    x, y = xyz[:, 0], xyz[:, 1]
    s = np.zeros((xyz.shape[0], 6))
    s[:, 0] = EXX * x + 0.5 * GXY * y + 0.02
    s[:, 1] = EYY * y + 0.5 * GXY * x - 0.01
    s[:, 2] = 0.5 * (KX * x * x + KY * y * y + 2 * KXY * x * y) + 0.3 + 0.1 * x - 0.2 * y
    s[:, 3] = KY * y + KXY * x - 0.2
    s[:, 4] = -(KX * x + KXY * y + 0.1)
    return s


def rotate_to_local(v, c2, s2, cs):
    """components (xx, yy, xy) of a symmetric 2-tensor in axes turned by the angle with cos^2, sin^2, cos*sin"""
    xx, yy, xy = v
    return np.array([c2 * xx + s2 * yy + 2 * cs * xy, s2 * xx + c2 * yy - 2 * cs * xy, -cs * xx + cs * yy + (c2 - s2) * xy])


@pytest.mark.parametrize("kind", ["q", "t"])
def test_constant_strain_and_curvature_patch(fso, kind):
    mesh, _ = fso.meshgen(kind, 5, 3, 0.0, 0.0, 10.0, 3.0, (1, 1, 1, 1), 1.0, 2, 1)
    out = fso.recover_resultants(mesh, patch_field(mesh.xyz), NU, E, T, quirks=0)
    Dm, Dp = material()
    sig = Dm @ np.array([EXX, EYY, GXY])
    # the Specht triangle's B yields +d2w, the DKQ's B yields -d2w (Batoz' beta = -grad w): M = Dp B w literally
    sign = 1.0 if kind == "t" else -1.0
    mom = sign * (Dp @ np.array([KX, KY, 2 * KXY]))
    for e in range(mesh.n_elem):
        en = mesh.enodes[mesh.eptr[e]:mesh.eptr[e + 1]]
        X = mesh.xyz[en]
        ux = (X[1] - X[0]) if kind == "t" else (0.5 * (X[1] + X[2]) - 0.5 * (X[3] + X[0]))   # fs.cpp:318 / :364
        ux = ux / np.linalg.norm(ux)
        c, s = ux[0], ux[1]
        want = np.concatenate([rotate_to_local(sig, c * c, s * s, c * s), rotate_to_local(mom, c * c, s * s, c * s)])
        assert np.allclose(out[e], want, rtol=1e-9, atol=1e-9 * np.abs(want).max()), (e, out[e], want)


def test_rigid_body_motion_gives_no_resultants(fso):
    from meshes import folded_cantilever
    m = folded_cantilever(skew=0.35)
    mesh = fso.Mesh(m["xyz"], m["etype"], m["eptr"], m["enodes"], m["bc"])
    om = np.array([0.01, -0.02, 0.015])
    s = np.zeros((mesh.n_nodes, 6))
    s[:, :3] = np.array([0.1, 0.2, -0.3]) + np.cross(om, mesh.xyz)
    s[:, 3:] = om
    for quirks in (0, 3):
        out = fso.recover_resultants(mesh, s, NU, E, T, quirks=quirks)
        scale = E * 0.01   # stress of a 1 % strain
        assert np.abs(out[:, :3]).max() < 1e-9 * scale
        assert np.abs(out[:, 3:]).max() < 1e-9 * scale * T ** 2

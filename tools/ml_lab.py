"""CPU lab (scipy) for the multilevel preconditioner design: iteration counts of PCG on oracle-assembled
plates with (a) Jacobi, (b) the additive hat-weighted lattice scheme, (c) a smoothed-aggregation V-cycle on
nested lattices with rigid-body modes.  Design evidence only; nothing in the product imports this."""
import sys, os, time
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from oracle import fso


def system(kind, n, thickness=0.5, L=10.0):
    mesh, forces = fso.meshgen(kind, n, n, 0, 0, L, L, (1, 1, 1, 1), 1.0, 2, 1, "z")
    S = fso.assemble(mesh, forces, 0.3, 1e7, thickness, threads=8)
    A = S.scipy().tocsr()
    xyz = np.zeros((S.n_dofnodes, 3))
    xyz[S.dofnode] = mesh.xyz
    mask = np.zeros(S.n_dofnodes, np.uint8)
    mask[S.dofnode] = S.mask
    return A, S.rhs.copy(), xyz, mask


def pcg(A, b, M, rtol=1e-8, maxit=200000):
    x = np.zeros_like(b)
    r = b.copy()
    z = M(r)
    p = z.copy()
    rz = r @ z
    bn = np.linalg.norm(b)
    for it in range(1, maxit + 1):
        q = A @ p
        a = rz / (p @ q)
        x += a * p
        r -= a * q
        if np.linalg.norm(r) <= rtol * bn:
            return x, it
        z = M(r)
        rz2 = r @ z
        p = z + (rz2 / rz) * p
        rz = rz2
    return x, maxit


def rbm_block(rho):
    """6x6: nodal (u,theta) of rigid-body mode coefficients (t, w) about a centre at -rho: u = t + w x rho"""
    B = np.zeros(rho.shape[:-1] + (6, 6))
    for a in range(6):
        B[..., a, a] = 1.0
    x, y, z = rho[..., 0], rho[..., 1], rho[..., 2]
    # u = t + w x rho  -> u_x = t_x + w_y z - w_z y ; u_y = t_y + w_z x - w_x z ; u_z = t_z + w_x y - w_y x
    B[..., 0, 4] = z; B[..., 0, 5] = -y
    B[..., 1, 5] = x; B[..., 1, 3] = -z
    B[..., 2, 3] = y; B[..., 2, 4] = -x
    return B


def tentative(xyz, mask, lo, H, npd, active):
    """aggregates = lattice cells of spacing H; returns BSR-ish P (6n x 6m) with RBMs about the cell centres"""
    n = xyz.shape[0]
    k = np.zeros((n, 3), np.int64)
    for d in range(3):
        if active[d]:
            k[:, d] = np.clip(np.floor((xyz[:, d] - lo[d]) / H).astype(np.int64), 0, npd[d] - 1)
    agg = (k[:, 2] * npd[1] + k[:, 1]) * npd[0] + k[:, 0]
    cen = np.array(lo)[None, :] + (k + 0.5) * H * np.array(active)[None, :]
    B = rbm_block(xyz - cen)
    free = np.array([[(m >> a) & 1 == 0 for a in range(6)] for m in mask], float)
    B = B * free[:, :, None]
    m = int(np.prod(npd))
    rows = (6 * np.arange(n)[:, None, None] + np.arange(6)[None, :, None] + 0 * np.arange(6)[None, None, :]).ravel()
    cols = (6 * agg[:, None, None] + 0 * np.arange(6)[None, :, None] + np.arange(6)[None, None, :]).ravel()
    P = sp.csr_matrix((B.ravel(), (rows, cols)), shape=(6 * n, 6 * m))
    # coarse "coordinates" and masks (coarse points never carry Dirichlet bits: handled by pseudo-inverse)
    kk = np.stack(np.meshgrid(np.arange(npd[0]), np.arange(npd[1]), np.arange(npd[2]), indexing="ij"), -1).reshape(-1, 3)
    order = (kk[:, 2] * npd[1] + kk[:, 1]) * npd[0] + kk[:, 0]
    cxyz = np.zeros((m, 3))
    cxyz[order] = np.array(lo)[None, :] + (kk + 0.5) * H * np.array(active)[None, :]
    return P, cxyz


def block_diag_pinv(A, nb):
    """pseudo-inverse of the 6x6 diagonal blocks as a block-diagonal sparse matrix"""
    Ab = sp.bsr_matrix(A, blocksize=(6, 6))
    Ab.sort_indices()
    rows = np.repeat(np.arange(nb), np.diff(Ab.indptr))
    sel = np.nonzero(Ab.indices == rows)[0]
    D = np.zeros((nb, 6, 6))
    D[rows[sel]] = Ab.data[sel]
    S = 0.5 * (D + D.transpose(0, 2, 1))
    w, V = np.linalg.eigh(S)
    wm = w.max(1, keepdims=True)
    ok = (wm > 0) & (w > 1e-12 * wm)
    inv = np.where(ok, 1.0 / np.where(ok, w, 1.0), 0.0)
    Di = np.einsum("nik,nk,njk->nij", V, inv, V)
    return sp.bsr_matrix((Di, np.arange(nb), np.arange(nb + 1)), shape=(6 * nb, 6 * nb)).tocsr()


def lam_max(DinvA, n, its=30):
    rng = np.random.default_rng(0)
    v = rng.standard_normal(n)
    lam = 1.0
    for _ in range(its):
        w = DinvA @ v
        lam = np.linalg.norm(w) / np.linalg.norm(v)
        v = w / np.linalg.norm(w)
    return lam


class SA:
    def __init__(self, A, xyz, mask, h, coarsen0=3, coarsen=3, block_smoother=True, omega_p=4.0 / 3.0, nu=1, min_pts=4, smooth_p=True,
                 cheb=0, gamma=1, fine_point=False):
        self.levels = []
        lo = xyz.min(0)
        ext = xyz.max(0) - lo
        active = [1 if e > 1e-9 * ext.max() else 0 for e in ext]
        H = coarsen0 * h
        self.nu = nu
        self.cheb = cheb
        self.gamma = gamma
        while True:
            n = A.shape[0] // 6
            Dinv = block_diag_pinv(A, n) if (block_smoother and not (fine_point and not self.levels)) else sp.diags(np.where(A.diagonal() != 0, 1.0 / np.where(A.diagonal() == 0, 1, A.diagonal()), 0))
            lam = 1.1 * lam_max(Dinv @ A, A.shape[0])
            lev = dict(A=A, Dinv=Dinv, lam=lam)
            self.levels.append(lev)
            npd = [max(1, int(np.ceil(ext[d] / H - 1e-9))) if active[d] else 1 for d in range(3)]
            if n <= min_pts * min_pts or len(self.levels) > 10:
                lev["dense"] = np.linalg.pinv(A.toarray(), rcond=1e-12, hermitian=True)
                break
            Pt, cxyz = tentative(xyz, mask, lo, H, npd, active)
            P = Pt - (omega_p / lam) * (Dinv @ (A @ Pt)) if smooth_p else Pt
            lev["P"] = P.tocsr()
            A = (P.T @ A @ P).tocsr()
            xyz, mask = cxyz, np.zeros(cxyz.shape[0], np.uint8)
            H *= coarsen
        print("  SA levels:", [l["A"].shape[0] // 6 for l in self.levels], "lam:", ["%.2f" % l["lam"] for l in self.levels])

    def smooth(self, lev, x, b):
        om = 1.0 / lev["lam"] * (4.0 / 3.0)
        if x is None:
            x = om * (lev["Dinv"] @ b)
            k0 = 1
        else:
            k0 = 0
        for _ in range(k0, self.nu):
            x = x + om * (lev["Dinv"] @ (b - lev["A"] @ x))
        return x

    def cycle(self, l, b):
        lev = self.levels[l]
        if "dense" in lev:
            return lev["dense"] @ b
        x = self.smooth(lev, None, b)
        for _ in range(self.gamma if l >= 1 else 1):
            r = b - lev["A"] @ x
            x = x + lev["P"] @ self.cycle(l + 1, lev["P"].T @ r)
        x = self.smooth(lev, x, b)
        return x

    def __call__(self, r):
        return self.cycle(0, r)


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "q"
    sizes = [int(s) for s in sys.argv[2].split(",")] if len(sys.argv) > 2 else [32, 64]
    thick = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
    for n in sizes:
        A, b, xyz, mask = system(kind, n, thick)
        h = 10.0 / n
        print("n=%d dof=%d" % (n, A.shape[0]))
        d = A.diagonal()
        t0 = time.time()
        _, itj = pcg(A, b, lambda r: r / d)
        print("  jacobi its", itj, "%.1fs" % (time.time() - t0))
        for kw in (dict(), dict(nu=2), dict(coarsen0=2, coarsen=3), dict(block_smoother=False), dict(smooth_p=False)):
            t0 = time.time()
            M = SA(A, xyz, mask, h, **kw)
            x, its = pcg(A, b, M)
            print("  SA", kw, "its", its, "%.1fs" % (time.time() - t0))



class HatBPX:
    """the additive scheme of the first draft (fs_mlpc.cuh before the V-cycle): Jacobi + sum_l P_l blockdiag(P_l^T A P_l)^+ P_l^T"""
    def __init__(self, A, xyz, mask, h, max_points=1 << 18):
        lo = xyz.min(0); ext = xyz.max(0) - lo; mx = ext.max()
        active = [1 if e > 1e-9 * mx else 0 for e in ext]
        n = xyz.shape[0]
        cap = max(27, min(max_points, n // 8))
        kfin = 1
        for k in range(2, 13):
            tot = 1
            for d in range(3):
                if active[d]: tot *= max(1, int(np.ceil(ext[d] / (mx / 2 ** k) - 1e-9))) + 1
            if tot > cap: break
            kfin = k
        free = np.array([[(m >> a) & 1 == 0 for a in range(6)] for m in mask], float)
        self.d = A.diagonal(); self.ops = []
        for l in range(kfin):
            H = mx / 2 ** (kfin - l)
            npd = [max(1, int(np.ceil(ext[d] / H - 1e-9))) + 1 if active[d] else 1 for d in range(3)]
            f = (xyz - lo) / H
            k = np.zeros((n, 3), np.int64); t = np.zeros((n, 3))
            for d in range(3):
                if active[d]:
                    k[:, d] = np.clip(f[:, d].astype(np.int64), 0, npd[d] - 2); t[:, d] = f[:, d] - k[:, d]
            rows, cols, vals = [], [], []
            for q in range(8):
                c = np.array([(q >> i) & 1 for i in range(3)])
                if any(c[d] and not active[d] for d in range(3)): continue
                w = np.ones(n)
                for d in range(3):
                    if active[d]: w *= t[:, d] if c[d] else 1 - t[:, d]
                ka = k + c
                a = (ka[:, 2] * npd[1] + ka[:, 1]) * npd[0] + ka[:, 0]
                cen = lo + ka * H
                B = rbm_block(xyz - cen) * free[:, :, None] * w[:, None, None]
                rows.append((6 * np.arange(n)[:, None, None] + np.arange(6)[None, :, None] + 0 * np.arange(6)[None, None, :]).ravel())
                cols.append((6 * a[:, None, None] + 0 * np.arange(6)[None, :, None] + np.arange(6)[None, None, :]).ravel())
                vals.append(B.ravel())
            m = int(np.prod(npd))
            P = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(6 * n, 6 * m))
            Ac = (P.T @ A @ P).tocsr()
            self.ops.append((P, block_diag_pinv(Ac, m)))
        print("  hat levels:", [o[1].shape[0] // 6 for o in self.ops])

    def __call__(self, r):
        z = r / self.d
        for P, Di in self.ops:
            z = z + P @ (Di @ (P.T @ r))
        return z


if __name__ == "__main__":
    main()

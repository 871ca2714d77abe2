"""scipy mirror of the multilevel preconditioner FS_PC_MLRBM (fem_shell_b200/csrc/fs_mlpc.cu), built with
explicit sparse matrices: tentative prolongators from rigid-body modes on lattice cells, damped block-Jacobi
prolongator smoothing, Galerkin products, V/W-cycle.  Test infrastructure: the GPU path builds the same
operators matrix-free (probing), so z = M^-1 r must agree to rounding when both use the same smoother weights."""
import numpy as np
import scipy.sparse as sp


def rbm_block(rho):
    B = np.zeros(rho.shape[:-1] + (6, 6))
    for a in range(6):
        B[..., a, a] = 1.0
    x, y, z = rho[..., 0], rho[..., 1], rho[..., 2]
    B[..., 0, 4] = z; B[..., 0, 5] = -y
    B[..., 1, 5] = x; B[..., 1, 3] = -z
    B[..., 2, 3] = y; B[..., 2, 4] = -x
    return B


def block_rows_to_csr(B, agg, m):
    n = B.shape[0]
    rows = (6 * np.arange(n)[:, None, None] + np.arange(6)[None, :, None] + 0 * np.arange(6)[None, None, :]).ravel()
    cols = (6 * agg[:, None, None] + 0 * np.arange(6)[None, :, None] + np.arange(6)[None, None, :]).ravel()
    return sp.csr_matrix((B.ravel(), (rows, cols)), shape=(6 * n, 6 * m))


def diag_blocks(A, nb):
    Ab = sp.bsr_matrix(A, blocksize=(6, 6))
    Ab.sort_indices()
    rows = np.repeat(np.arange(nb), np.diff(Ab.indptr))
    sel = np.nonzero(Ab.indices == rows)[0]
    D = np.zeros((nb, 6, 6))
    D[rows[sel]] = Ab.data[sel]
    return D


def block_diag(D):
    nb = D.shape[0]
    return sp.bsr_matrix((D, np.arange(nb), np.arange(nb + 1)), shape=(6 * nb, 6 * nb)).tocsr()


def pinv_blocks(D):
    S = 0.5 * (D + D.transpose(0, 2, 1))
    w, V = np.linalg.eigh(S)
    wm = w.max(1, keepdims=True)
    ok = (wm > 0) & (w > 1e-12 * wm)
    inv = np.where(ok, 1.0 / np.where(ok, w, 1.0), 0.0)
    return np.einsum("nik,nk,njk->nij", V, inv, V)


def element_extent(xyz_nodes, eptr, enodes):
    h = np.zeros(3)
    for e in range(len(eptr) - 1):
        P = xyz_nodes[enodes[eptr[e]:eptr[e + 1]]]
        h = np.maximum(h, P.max(0) - P.min(0))
    return h


class Mirror:
    def __init__(self, A, xyz, mask, h, cells, lambdas, gamma=2, scale=1.0):
        """A: scipy CSR in dof order; xyz, mask: per dof-node; h: element extent per axis; cells: list of
        (nx, ny, nz) per lattice level and lambdas: [mesh, lattice 0, ...] as reported by FemShell.ml_info()"""
        self.gamma = gamma
        lo_box, hi_box = xyz.min(0), xyz.max(0)
        ext = hi_box - lo_box
        active = np.array([1 if e > 1e-9 * ext.max() else 0 for e in ext])
        H = np.where(active == 1, 3.0 * np.where(h > 0, h, ext) * scale, 0.0)
        lo = np.where(active == 1, lo_box - 0.5 * np.where(h > 0, h, ext), lo_box)
        np0 = [int(np.floor((hi_box[d] - lo[d]) / H[d])) + 1 if active[d] else 1 for d in range(3)]
        assert tuple(np0) == tuple(cells[0]), (np0, cells[0])
        free = np.array([[0.0 if (m >> a) & 1 else 1.0 for a in range(6)] for m in mask])
        self.levels = []
        n = xyz.shape[0]
        D = diag_blocks(A, n)
        Dinv = block_diag(np.linalg.inv(D))
        om = (4.0 / 3.0) / lambdas[0]
        # mesh -> lattice 0
        k = np.zeros((n, 3), np.int64)
        for d in range(3):
            if active[d]:
                k[:, d] = np.clip(((xyz[:, d] - lo[d]) / H[d]).astype(np.int64), 0, np0[d] - 1)
        agg = (k[:, 2] * np0[1] + k[:, 1]) * np0[0] + k[:, 0]
        cen = lo[None, :] + (k + 0.5) * H[None, :]
        Pt = block_rows_to_csr(rbm_block(xyz - cen) * free[:, :, None], agg, int(np.prod(np0)))
        P = (Pt - om * (Dinv @ (A @ Pt))).tocsr()
        self.levels.append(dict(A=A, Dinv=Dinv, om=om, P=P))
        Ac = (P.T @ A @ P).tocsr()
        npd = list(np0)
        for l in range(len(cells)):
            m = int(np.prod(npd))
            if l == len(cells) - 1:
                Md = Ac.toarray()
                Md = 0.5 * (Md + Md.T)
                dm = Md.diagonal().max()
                keep = np.nonzero(Md.diagonal() > 1e-12 * dm)[0]
                inv = np.zeros_like(Md)
                inv[np.ix_(keep, keep)] = np.linalg.inv(Md[np.ix_(keep, keep)])
                self.levels.append(dict(dense=inv))
                break
            Dl = pinv_blocks(diag_blocks(Ac, m))
            Dinv = block_diag(Dl)
            om = (4.0 / 3.0) / lambdas[1 + l]
            npn = [(npd[d] + 2) // 3 if active[d] else 1 for d in range(3)]
            assert tuple(npn) == tuple(cells[l + 1])
            kk = np.stack(np.meshgrid(np.arange(npd[0]), np.arange(npd[1]), np.arange(npd[2]), indexing="ij"), -1).reshape(-1, 3)
            idx = (kk[:, 2] * npd[1] + kk[:, 1]) * npd[0] + kk[:, 0]
            order = np.argsort(idx)
            kk = kk[order]
            kp = np.where(active[None, :] == 1, kk // 3, 0)
            agg = (kp[:, 2] * npn[1] + kp[:, 1]) * npn[0] + kp[:, 0]
            cc = lo[None, :] + (kk + 0.5) * H[None, :]
            cp = lo[None, :] + (kp + 0.5) * (3.0 * H)[None, :]
            Pt = block_rows_to_csr(rbm_block(cc - cp), agg, int(np.prod(npn)))
            P = (Pt - om * (Dinv @ (Ac @ Pt))).tocsr()
            self.levels.append(dict(A=Ac, Dinv=Dinv, om=om, P=P))
            Ac = (P.T @ Ac @ P).tocsr()
            npd, H = npn, 3.0 * H

    def lambda_max(self, l, its=200):
        lev = self.levels[l]
        v = np.random.default_rng(l).standard_normal(lev["A"].shape[0])
        lam = 0.0
        for _ in range(its):
            w = lev["Dinv"] @ (lev["A"] @ v)
            lam = np.linalg.norm(w) / np.linalg.norm(v)
            v = w / np.linalg.norm(w)
        return lam

    def cycle(self, l, b):
        lev = self.levels[l]
        if "dense" in lev:
            return lev["dense"] @ b
        A, Dinv, om, P = lev["A"], lev["Dinv"], lev["om"], lev["P"]
        x = om * (Dinv @ b)
        digits = [int(ch) for ch in str(self.gamma)]          # digit k = visits of lattice k+1 per visit of lattice k
        g = digits[min(l - 1, len(digits) - 1)] if l >= 1 else 1
        if "dense" in self.levels[l + 1]:
            g = 1                     # exact coarse solve: a second visit computes a zero correction (fs_mlpc.cu lat_cycle)
        for _ in range(g if l >= 1 else 1):
            x = x + P @ self.cycle(l + 1, P.T @ (b - A @ x))
        return x + om * (Dinv @ (b - A @ x))

    def __call__(self, r):
        return self.cycle(0, r)


def stencil_to_csr(Aarr, cells, active):
    """lattice stencil as exported by fs_debug_ml_level(what=0) -> scipy CSR (6n x 6n)"""
    npd = list(cells)
    n = int(np.prod(npd))
    n6 = 6 * n
    nact = int(sum(active))
    ns = 3 ** nact
    A = Aarr.reshape(ns, 6, n, 6)            # [s][b][p][a]
    kk = np.stack(np.meshgrid(np.arange(npd[0]), np.arange(npd[1]), np.arange(npd[2]), indexing="ij"), -1).reshape(-1, 3)
    idx = (kk[:, 2] * npd[1] + kk[:, 1]) * npd[0] + kk[:, 0]
    kk = kk[np.argsort(idx)]
    rows, cols, vals = [], [], []
    for s in range(ns):
        o = np.zeros(3, np.int64)
        ss = s
        for d in range(3):
            if active[d]:
                o[d] = ss % 3 - 1
                ss //= 3
        nb = kk + o
        ok = np.all((nb >= 0) & (nb < np.array(npd)), axis=1)
        p = np.nonzero(ok)[0]
        q = (nb[p, 2] * npd[1] + nb[p, 1]) * npd[0] + nb[p, 0]
        blk = A[s][:, p, :].transpose(1, 2, 0)     # [p][a][b]
        rows.append((6 * p[:, None, None] + np.arange(6)[None, :, None] + 0 * np.arange(6)[None, None, :]).ravel())
        cols.append((6 * q[:, None, None] + 0 * np.arange(6)[None, :, None] + np.arange(6)[None, None, :]).ravel())
        vals.append(blk.ravel())
    return sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n6, n6))

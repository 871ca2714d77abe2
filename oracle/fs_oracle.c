/*
 * fs_oracle.c -- CPU restatement of fem-shell's assembly + solve hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, the smoke
 * check in __graft_entry__.py and bench.py's cpu_baseline / --impl reference
 * legs may load it.  The product (fem_shell_b200/csrc) never links, calls or
 * falls back to anything in this directory.
 *
 * What it restates (reference = precice/fem-shell, src/fem-shell/fem-shell.cpp,
 * abbreviated "fs.cpp"):
 *   material matrices            fs.cpp:273-294
 *   element frame / geometry     fs.cpp:306-433
 *   membrane stiffness           fs.cpp:443-542
 *   plate stiffness              fs.cpp:551-688, evalBTri :698-891, evalBQuad :901-990
 *   shell superposition+drilling fs.cpp:999-1053
 *   local->global + permutation  fs.cpp:1061-1110
 *   element RHS                  fs.cpp:1118-1153
 *   element loop / constraint / accumulate   fs.cpp:1160-1233
 * plus the libMesh / PETSc semantics those lines rely on and which are NOT in
 * the reference tree (libMesh master ~Dec 2015, PETSc maint ~3.6, both
 * unpinned by the reference README): DenseMatrix multiply order, the
 * non-const DenseMatrix::det() (in-place LU, stale flag), DofMap numbering,
 * Dirichlet element constraint, dense 6x6 node-block sparsity, Krylov solve.
 *
 * Parity status: end-to-end displacements are PINNED against the thesis
 * tables (doc/validation.tex, Tests A-G) in tests/test_oracle_goldens.py.
 * Individual stiffness entries, CSR ordering and the det() side effects on
 * non-rectangular quads are "parity unpinned": no reference output exists for
 * them in this container (libMesh/PETSc cannot be built here).
 * fso_recover_resultants restates formulas the thesis prints but the reference
 * never coded (doc/shellelements.tex:524, :1394-1403); it is pinned to
 * closed-form constant-strain / constant-curvature / rigid-body fields in
 * tests/test_oracle_resultants.py.
 *
 * Plain C99, doubles, element-id order, no reassociation tricks.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define FSO_TRI3 3  /* XDA element type id, src/meshgen/main_all.cpp:245 */
#define FSO_QUAD4 5 /* src/meshgen/main_all.cpp:247 */

/* quirk switches */
#define FSO_QUIRK_Y21 1      /* fs.cpp:586 uses x31*x31 */
#define FSO_QUIRK_DET_LU 2   /* DenseMatrix::det() leaves LU factors behind, fs.cpp:512,652 */
#define FSO_QUIRKS_REFERENCE (FSO_QUIRK_Y21 | FSO_QUIRK_DET_LU)

/* ------------------------------------------------------------------ */
/* tiny dense helpers with libMesh DenseMatrix semantics (row-major)  */
/* ------------------------------------------------------------------ */

/* C(m x n) = A(m x k) * B(k x n); used for right_multiply / left_multiply */
static void mm(int m, int k, int n, const double *A, const double *B, double *C)
{
    for (int i = 0; i < m; i++)
        for (int j = 0; j < n; j++) {
            double s = 0.0;
            for (int l = 0; l < k; l++)
                s += A[i * k + l] * B[l * n + j];
            C[i * n + j] = s;
        }
}

/* C(k x n) = A^T * B with A (m x k), B (m x n): left_multiply_transpose */
static void mtm(int m, int k, int n, const double *A, const double *B, double *C)
{
    for (int i = 0; i < k; i++)
        for (int j = 0; j < n; j++) {
            double s = 0.0;
            for (int l = 0; l < m; l++)
                s += A[l * k + i] * B[l * n + j];
            C[i * n + j] = s;
        }
}

/*
 * libMesh DenseMatrix<T>::det() on a 2x2, restated: if the matrix has not been
 * factorised (*lu_flag == 0) run the in-place LU with partial pivoting (swap
 * only on strictly larger magnitude, scale the upper row by 1/diag, Schur
 * update), remember the pivot rows and raise the flag.  Then return the
 * product of the CURRENT diagonal, sign flipped once per recorded row swap.
 */
static double det2_libmesh(double J[4], int *lu_flag, int piv[2])
{
    if (!*lu_flag) {
        piv[0] = 0;
        if (fabs(J[0]) < fabs(J[2])) {
            piv[0] = 1;
            double t0 = J[0], t1 = J[1];
            J[0] = J[2]; J[1] = J[3];
            J[2] = t0;   J[3] = t1;
        }
        double dinv = 1.0 / J[0];
        J[1] *= dinv;
        J[3] -= J[2] * J[1];
        piv[1] = 1;
        *lu_flag = 1;
    }
    double d = 1.0;
    if (piv[0] != 0) d *= -1.0;
    d *= J[0];
    if (piv[1] != 1) d *= -1.0;
    d *= J[3];
    return d;
}

/* ------------------------------------------------------------------ */
/* fs.cpp:273-294  initMaterialMatrices                                */
/* ------------------------------------------------------------------ */
void fso_material(double nu, double em, double thickness, double *Dm, double *Dp)
{
    double D[9] = {1.0, nu, 0.0, nu, 1.0, 0.0, 0.0, 0.0, (1.0 - nu) / 2.0};
    double fm = em / (1.0 - nu * nu);
    double fp = em * pow(thickness, 3.0) / (12.0 * (1.0 - nu * nu));
    for (int i = 0; i < 9; i++) {
        Dm[i] = D[i] * fm;
        Dp[i] = D[i] * fp;
    }
}

/* ------------------------------------------------------------------ */
/* fs.cpp:306-433  initElement                                         */
/* xyz: nen x 3 global node coordinates.  Outputs: trafo (3x3, rows = */
/* local axes), loc (3 x nen local coords; TRI3 uses columns 0,1 for  */
/* nodes B,C with A at the origin), dphi (nen x 2), area.              */
/* ------------------------------------------------------------------ */
static void cross3(const double *a, const double *b, double *c)
{
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = -a[0] * b[2] + a[2] * b[0];
    c[2] = a[0] * b[1] - a[1] * b[0];
}
static double norm3(const double *a) { return sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }
static void unit3(double *a)
{
    double l = norm3(a);
    a[0] /= l; a[1] /= l; a[2] /= l;
}

static void init_element(int type, const double *xyz, double *trafo, double *loc, double *dphi,
                         double *area)
{
    double U[3], V[3], W[3];
    double glob[12];
    int ncol;
    if (type == FSO_TRI3) {
        /* fs.cpp:318-340 */
        for (int i = 0; i < 3; i++) {
            U[i] = xyz[3 + i] - xyz[i];
            V[i] = xyz[6 + i] - xyz[i];
        }
        ncol = 2;
        for (int i = 0; i < 3; i++) {
            glob[i * 2 + 0] = U[i];
            glob[i * 2 + 1] = V[i];
        }
        cross3(U, V, W);
        *area = 0.5 * norm3(W);
        unit3(U);
        unit3(W);
        cross3(W, U, V);
    } else {
        /* fs.cpp:345-375 : edge midpoints, x axis through J-L */
        const double *A = xyz, *B = xyz + 3, *C = xyz + 6, *D = xyz + 9;
        double nI[3], nJ[3], nK[3], nL[3], R[3];
        for (int i = 0; i < 3; i++) {
            nI[i] = A[i] + 0.5 * (B[i] - A[i]);
            nJ[i] = B[i] + 0.5 * (C[i] - B[i]);
            nK[i] = C[i] + 0.5 * (D[i] - C[i]);
            nL[i] = D[i] + 0.5 * (A[i] - D[i]);
        }
        ncol = 4;
        for (int n = 0; n < 4; n++)
            for (int i = 0; i < 3; i++)
                glob[i * 4 + n] = xyz[3 * n + i];
        for (int i = 0; i < 3; i++) {
            U[i] = nJ[i] - nL[i];
            R[i] = nK[i] - nI[i];
        }
        unit3(U);
        cross3(U, R, W);
        unit3(W);
        cross3(W, U, V);
    }
    /* fs.cpp:378-384 */
    for (int i = 0; i < 3; i++) {
        trafo[0 + i] = U[i];
        trafo[3 + i] = V[i];
        trafo[6 + i] = W[i];
    }
    /* fs.cpp:391 transUV.left_multiply(trafo) */
    mm(3, 3, ncol, trafo, glob, loc);

    if (type == FSO_TRI3) {
        /* fs.cpp:405-411 ; loc row 0 = x, row 1 = y ; col 0 = node B, col 1 = node C */
        dphi[0] = -loc[0];          /* x12 */
        dphi[2] = loc[1];           /* x31 */
        dphi[4] = loc[0] - loc[1];  /* x23 */
        dphi[1] = -loc[2];          /* y12 */
        dphi[3] = loc[3];           /* y31 */
        dphi[5] = loc[2] - loc[3];  /* y23 */
    } else {
        /* fs.cpp:415-431 */
        for (int i = 0; i < 4; i++) {
            int j = (i + 1) % 4;
            dphi[2 * i + 0] = loc[0 * 4 + i] - loc[0 * 4 + j];
            dphi[2 * i + 1] = loc[1 * 4 + i] - loc[1 * 4 + j];
        }
        double a = 0.0;
        for (int i = 0; i < 4; i++) {
            int j = (i + 1) % 4;
            a += loc[i] * loc[4 + j] - loc[j] * loc[4 + i];
        }
        *area = a * 0.5;
    }
}

/* ------------------------------------------------------------------ */
/* fs.cpp:443-542  calcPlane                                           */
/* ------------------------------------------------------------------ */
static void calc_plane(int type, const double *loc, const double *dphi, double area,
                       const double *Dm, double thickness, int quirks, double *Kem)
{
    if (type == FSO_TRI3) {
        /* fs.cpp:448-467 */
        double B[18];
        memset(B, 0, sizeof B);
        double x12 = dphi[0], y12 = dphi[1], x31 = dphi[2], y31 = dphi[3], x23 = dphi[4],
               y23 = dphi[5];
        B[0 * 6 + 0] = y23;  B[0 * 6 + 2] = y31;  B[0 * 6 + 4] = y12;
        B[1 * 6 + 1] = -x23; B[1 * 6 + 3] = -x31; B[1 * 6 + 5] = -x12;
        B[2 * 6 + 0] = -x23; B[2 * 6 + 1] = y23;
        B[2 * 6 + 2] = -x31; B[2 * 6 + 3] = y31;
        B[2 * 6 + 4] = -x12; B[2 * 6 + 5] = y12;
        double s = 1.0 / (2.0 * area);
        for (int i = 0; i < 18; i++) B[i] *= s;
        double DB[18];
        mm(3, 3, 6, Dm, B, DB);
        mtm(3, 6, 6, B, DB, Kem);
        double f = thickness * area;
        for (int i = 0; i < 36; i++) Kem[i] *= f;
        return;
    }
    /* QUAD4: fs.cpp:472-540 */
    double root = sqrt(1.0 / 3.0);
    memset(Kem, 0, 64 * sizeof(double));
    double G[32];
    memset(G, 0, sizeof G);
    for (int ii = 0; ii < 2; ii++) {
        double r = pow(-1.0, ii) * root;
        for (int jj = 0; jj < 2; jj++) {
            double s = pow(-1.0, jj) * root;
            double dhdr[4], dhds[4];
            dhdr[0] = -0.25 * (1 - s); dhdr[1] = 0.25 * (1 - s);
            dhdr[2] = 0.25 * (1 + s);  dhdr[3] = -0.25 * (1 + s);
            dhds[0] = -0.25 * (1 - r); dhds[1] = -0.25 * (1 + r);
            dhds[2] = 0.25 * (1 + r);  dhds[3] = 0.25 * (1 - r);
            double J[4] = {0, 0, 0, 0};
            for (int i = 0; i < 4; i++) {
                J[0] += dhdr[i] * loc[0 * 4 + i];
                J[1] += dhdr[i] * loc[1 * 4 + i];
                J[2] += dhds[i] * loc[0 * 4 + i];
                J[3] += dhds[i] * loc[1 * 4 + i];
            }
            double det;
            if (quirks & FSO_QUIRK_DET_LU) {
                /* fs.cpp:504 resize() clears the LU flag, so every GP factorises;
                 * fs.cpp:515-517 then read the overwritten J */
                int flag = 0, piv[2];
                det = det2_libmesh(J, &flag, piv);
            } else {
                det = J[0] * J[3] - J[1] * J[2];
            }
            double Bm[12];
            memset(Bm, 0, sizeof Bm);
            Bm[0 * 4 + 0] = J[3];  Bm[0 * 4 + 1] = -J[1];
            Bm[1 * 4 + 2] = -J[2]; Bm[1 * 4 + 3] = J[0];
            Bm[2 * 4 + 0] = -J[2]; Bm[2 * 4 + 1] = J[0];
            Bm[2 * 4 + 2] = J[3];  Bm[2 * 4 + 3] = -J[1];
            double di = 1.0 / det;
            for (int i = 0; i < 12; i++) Bm[i] *= di;
            for (int i = 0; i < 4; i++) {
                G[0 * 8 + 2 * i] = dhdr[i];
                G[1 * 8 + 2 * i] = dhds[i];
                G[2 * 8 + 1 + 2 * i] = dhdr[i];
                G[3 * 8 + 1 + 2 * i] = dhds[i];
            }
            double B[24], BtD[24], tmp[64];
            mm(3, 4, 8, Bm, G, B);      /* fs.cpp:529 */
            mtm(3, 8, 3, B, Dm, BtD);   /* fs.cpp:534 : B^T Dm (8x3) */
            mm(8, 3, 8, BtD, B, tmp);   /* fs.cpp:535 */
            double f = det * thickness;
            for (int i = 0; i < 64; i++) Kem[i] += tmp[i] * f;
        }
    }
}

/* ------------------------------------------------------------------ */
/* fs.cpp:698-891  evalBTri                                            */
/* Second derivatives of Specht's nine shape functions in area         */
/* coordinates.  The reference spells every entry out; the repeated    */
/* brackets are named here (S*, R*, T*) but each bracket keeps the     */
/* reference's term order.                                             */
/* ------------------------------------------------------------------ */
static void eval_b_tri(const double *C, double L1, double L2, const double *dphi, double *out)
{
    double x12 = dphi[0], y12 = dphi[1], x31 = dphi[2], y31 = dphi[3], x23 = dphi[4],
           y23 = dphi[5];
    double mu1 = (C[0] - C[1]) / C[2];
    double mu2 = (C[2] - C[0]) / C[1];
    double mu3 = (C[1] - C[2]) / C[0];
    double L3 = 1 - L1 - L2;
    double p1 = 1 + 3 * mu1, p2 = 1 + 3 * mu2, p3 = 1 + 3 * mu3;
    double q3 = 1 - 3 * mu3, m2 = -1 + 3 * mu2, n3 = -1 - 3 * mu3;
    double a = 3 * (1 - mu3) * L1 - p3 * L2 + p3 * L3;
    double b = 3 * (1 - mu2) * L3 - p2 * L1 + p2 * L2;
    double c = 3 * (1 - mu1) * L2 - p1 * L3 + p1 * L1;

    /* row 0 (fs.cpp:725-747) */
    double S1 = -2 + 6 * L1 + 4 * L2 - L2 * b - 4 * L2 * L3 + 4 * L1 * L2;
    double S2 = 2 * L2 - L2 * a + L2 * L3 * 2 * q3 - L1 * L2 * 2 * q3;
    double S3 = -L2 * c + L2 * L3 * 2 * p1 - L1 * L2 * 2 * p1;
    double S4 = -4 + 6 * L1 + 4 * L2 - L2 * b - 4 * L2 * L3 + 4 * L1 * L2;
    out[0] = 6 + L2 * (-4 - 2 * a) + 4 * q3 * (L2 * L3 - L1 * L2) - 12 * L1 + 2 * L2 * b +
             8 * (L2 * L3 - L1 * L2);
    out[1] = -y31 * S1 - y12 * S2;
    out[2] = x31 * S1 + x12 * S2;
    out[3] = -2 * L2 * c + 4 * p1 * (L2 * L3 - L1 * L2) - 4 * L2 + 2 * L2 * a +
             4 * q3 * (-L2 * L3 + L1 * L2);
    out[4] = -y12 * S2 - y23 * S3;
    out[5] = x12 * S2 + x23 * S3;
    out[6] = -6 + 12 * L1 + 8 * L2 - 2 * L2 * b + 8 * (L1 * L2 - L2 * L3) + 2 * L2 * c +
             4 * p1 * (L1 * L2 - L2 * L3);
    out[7] = -y23 * S3 - y31 * S4;
    out[8] = x23 * S3 + x31 * S4;

    /* row 1 (fs.cpp:749-771) */
    double R1 = 2 * L1 - 1 * L1 * b + 1 * L1 * L3 * 2 * m2 - 1 * L1 * L2 * 2 * m2;
    double R2 = -1 * L1 * a + 1 * L1 * L3 * 2 * n3 - 1 * L1 * L2 * 2 * n3;
    double R3 = -6 * L2 + 2 - 2 * L1 - 1 * L1 * c + 4 * L3 * L1 - 4 * L1 * L2;
    double R4 = -6 * L2 + 4 - 2 * L1 - 1 * L1 * c + 4 * L3 * L1 - 4 * L1 * L2;
    out[9 + 0] = -2 * L1 * a + 2 * L1 * L3 * 2 * n3 - 2 * L1 * L2 * 2 * n3 - 4 * L1 + 2 * L1 * b -
                 2 * L1 * L3 * 2 * m2 + 2 * L1 * L2 * 2 * m2;
    out[9 + 1] = -y31 * R1 - y12 * R2;
    out[9 + 2] = x31 * R1 + x12 * R2;
    out[9 + 3] = 6 - 12 * L2 - 4 * L1 - 2 * L1 * c + 8 * L3 * L1 - 8 * L1 * L2 + 2 * L1 * a -
                 2 * L1 * L3 * 2 * n3 + 2 * L1 * L2 * 2 * n3;
    out[9 + 4] = -y12 * R2 - y23 * R3;
    out[9 + 5] = x12 * R2 + x23 * R3;
    out[9 + 6] = -6 + 8 * L1 - 2 * L1 * b + 2 * L1 * L3 * 2 * m2 - 2 * L1 * L2 * 2 * m2 + 12 * L2 +
                 2 * L1 * c - 8 * L3 * L1 + 8 * L1 * L2;
    out[9 + 7] = -y23 * R4 - y31 * R1;
    out[9 + 8] = x23 * R4 + x31 * R1;

    /* row 2 (fs.cpp:773-887), doubled at fs.cpp:889-890 */
    double T1 = -1 + 4 * L1 + 2 * L2 + 0.5 * L3 * b - 0.5 * L2 * b + 0.5 * L2 * L3 * 2 * m2 -
                0.5 * L1 * b - 0.5 * L1 * L2 * 2 * m2 - 2 * L3 * L1 + 2 * L1 * L2;
    double T2 = 2 * L1 + 0.5 * L3 * a - 0.5 * L2 * a + 0.5 * L2 * L3 * 2 * n3 - 0.5 * L1 * a -
                0.5 * L1 * L2 * 2 * n3 + 0.5 * L1 * L3 * 2 * q3 - 0.5 * L1 * L2 * 2 * q3;
    double T3 = T2 - 1;
    double T4 = -2 * L2 + 0.5 * L3 * c - 0.5 * L2 * c + 2 * L2 * L3 - 0.5 * L1 * c - 2 * L1 * L2 +
                0.5 * L1 * L3 * 2 * p1 - 0.5 * L1 * L2 * 2 * p1;
    double T5 = T4 + 1;
    double T6 = -2 + 4 * L1 + 2 * L2 + 0.5 * L3 * b - 0.5 * L2 * b + 0.5 * L2 * L3 * 2 * m2 -
                0.5 * L1 * b - 0.5 * L1 * L2 * 2 * m2 - 2 * L3 * L1 + 2 * L1 * L2;
    out[18 + 0] = 2 - 4 * L1 + L3 * a - L2 * a + L2 * L3 * 2 * n3 - L1 * a - L1 * L2 * 2 * n3 +
                  L1 * L3 * 2 * q3 - L1 * L2 * 2 * q3 - 4 * L2 - L3 * b + L2 * b -
                  L2 * L3 * 2 * m2 + L1 * b + L1 * L2 * 2 * m2 + 4 * L3 * L1 - 4 * L1 * L2;
    out[18 + 1] = -y31 * T1 - y12 * T2;
    out[18 + 2] = x31 * T1 + x12 * T2;
    out[18 + 3] = 2 - 4 * L2 + L3 * c - L2 * c + 4 * L2 * L3 - L1 * c - 4 * L1 * L2 +
                  L1 * L3 * 2 * p1 - L1 * L2 * 2 * p1 - 4 * L1 - L3 * a + L2 * a + L1 * a -
                  L2 * L3 * 2 * n3 + L1 * L2 * 2 * n3 - L1 * L3 * 2 * q3 + L1 * L2 * 2 * q3;
    out[18 + 4] = -y12 * T3 - y23 * T4;
    out[18 + 5] = x12 * T3 + x23 * T4;
    out[18 + 6] = -4 + 8 * L1 + 8 * L2 + L3 * b - L2 * b + L2 * L3 * 2 * m2 - L1 * b -
                  L1 * L2 * 2 * m2 - 4 * L3 * L1 + 8 * L1 * L2 - L3 * c + L2 * c - 4 * L2 * L3 +
                  L1 * c - L1 * L3 * 2 * p1 + L1 * L2 * 2 * p1;
    out[18 + 7] = -y23 * T5 - y31 * T6;
    out[18 + 8] = x23 * T5 + x31 * T6;
    for (int i = 0; i < 9; i++) out[18 + i] *= 2.0;
}

/* ------------------------------------------------------------------ */
/* fs.cpp:901-990  evalBQuad (DKQ)                                     */
/* H: 5 x 4 coefficients a..e for sides 5..8 (row-major).              */
/* ------------------------------------------------------------------ */
static void eval_b_quad(const double *H, double xi, double eta, const double *Jinv, double *out)
{
    double Nx[8], Ne[8];
    Nx[0] = 0.25 * (2.0 * xi + eta) * (1.0 - eta);
    Nx[1] = 0.25 * (2.0 * xi - eta) * (1.0 - eta);
    Nx[2] = 0.25 * (2.0 * xi + eta) * (1.0 + eta);
    Nx[3] = 0.25 * (2.0 * xi - eta) * (1.0 + eta);
    Nx[4] = -xi * (1.0 - eta);
    Nx[5] = 0.5 * (1.0 - pow(eta, 2.0));
    Nx[6] = -xi * (1.0 + eta);
    Nx[7] = -0.5 * (1.0 - pow(eta, 2.0));
    Ne[0] = 0.25 * (2.0 * eta + xi) * (1.0 - xi);
    Ne[1] = 0.25 * (2.0 * eta - xi) * (1.0 + xi);
    Ne[2] = 0.25 * (2.0 * eta + xi) * (1.0 + xi);
    Ne[3] = 0.25 * (2.0 * eta - xi) * (1.0 - xi);
    Ne[4] = -0.5 * (1.0 - pow(xi, 2.0));
    Ne[5] = -eta * (1.0 + xi);
    Ne[6] = 0.5 * (1.0 - pow(xi, 2.0));
    Ne[7] = -eta * (1.0 - xi);

    const double *ha = H, *hb = H + 4, *hc = H + 8, *hd = H + 12, *he = H + 16;
    double Hx[2][12], Hy[2][12];
    for (int d = 0; d < 2; d++) {
        const double *N = d ? Ne : Nx;
        /* node k (0..3) couples mid-side k (after it) and mid-side k-1 (before it):
         * fs.cpp:931-981 written as a loop over the four corner nodes */
        for (int k = 0; k < 4; k++) {
            int s = k, p = (k + 3) % 4; /* side index i5..i8 */
            double Ns = N[4 + s], Np = N[4 + p];
            Hx[d][3 * k + 0] = 1.5 * (ha[s] * Ns - ha[p] * Np);
            Hx[d][3 * k + 1] = hb[s] * Ns + hb[p] * Np;
            Hx[d][3 * k + 2] = N[k] - hc[s] * Ns - hc[p] * Np;
            Hy[d][3 * k + 0] = 1.5 * (hd[s] * Ns - hd[p] * Np);
            Hy[d][3 * k + 1] = -N[k] + he[s] * Ns + he[p] * Np;
            Hy[d][3 * k + 2] = -Hx[d][3 * k + 1];
        }
    }
    /* fs.cpp:984-989 */
    for (int i = 0; i < 12; i++) {
        out[i] = Jinv[0] * Hx[0][i] + Jinv[1] * Hx[1][i];
        out[12 + i] = Jinv[2] * Hy[0][i] + Jinv[3] * Hy[1][i];
        out[24 + i] = Jinv[0] * Hy[0][i] + Jinv[1] * Hy[1][i] + Jinv[2] * Hx[0][i] +
                      Jinv[3] * Hx[1][i];
    }
}

/* ------------------------------------------------------------------ */
/* fs.cpp:551-688  calcPlate                                           */
/* ------------------------------------------------------------------ */
static void calc_plate(int type, const double *dphi, double area, const double *Dp, int quirks,
                       double *Kep)
{
    if (type == FSO_TRI3) {
        static const double qps[3][2] = {
            {1.0 / 6.0, 1.0 / 6.0}, {2.0 / 3.0, 1.0 / 6.0}, {1.0 / 6.0, 2.0 / 3.0}};
        double side[3];
        for (int i = 0; i < 3; i++)
            side[i] = pow(dphi[2 * i], 2.0) + pow(dphi[2 * i + 1], 2.0);
        memset(Kep, 0, 81 * sizeof(double));
        double x31 = dphi[2], y31 = dphi[3], x23 = dphi[4], y23 = dphi[5];
        for (int g = 0; g < 3; g++) {
            double B[27];
            eval_b_tri(side, qps[g][0], qps[g][1], dphi, B);
            double Y[9];
            Y[0] = pow(y23, 2.0);
            Y[1] = pow(y31, 2.0);
            Y[2] = y23 * y31;
            Y[3] = pow(x23, 2.0);
            Y[4] = pow(x31, 2.0);
            Y[5] = x31 * x23;
            Y[6] = -2.0 * x23 * y23;
            if (quirks & FSO_QUIRK_Y21)
                Y[7] = -2.0 * x31 * x31; /* fs.cpp:586 as written */
            else
                Y[7] = -2.0 * x31 * y31; /* doc/shellelements.tex:1131 */
            Y[8] = -x23 * y31 - x31 * y23;
            double sc = 1.0 / (4.0 * pow(area, 2.0));
            for (int i = 0; i < 9; i++) Y[i] *= sc;
            double t1[9], t2[27], t3[27], t4[81];
            mm(3, 3, 3, Dp, Y, t1);    /* fs.cpp:592 */
            mm(3, 3, 9, t1, B, t2);    /* fs.cpp:593 */
            mtm(3, 3, 9, Y, t2, t3);   /* fs.cpp:594 */
            mtm(3, 9, 9, B, t3, t4);   /* fs.cpp:595 */
            for (int i = 0; i < 81; i++) Kep[i] += t4[i] * (1.0 / 6.0);
        }
        for (int i = 0; i < 81; i++) Kep[i] *= 2.0 * area;
        return;
    }
    /* QUAD4 / DKQ : fs.cpp:607-686 */
    double side[4], H[20];
    for (int i = 0; i < 4; i++)
        side[i] = pow(dphi[2 * i], 2.0) + pow(dphi[2 * i + 1], 2.0);
    for (int i = 0; i < 4; i++) {
        double dx = dphi[2 * i], dy = dphi[2 * i + 1];
        H[0 * 4 + i] = -dx / side[i];
        H[1 * 4 + i] = 0.75 * dx * dy / side[i];
        H[2 * 4 + i] = (0.25 * pow(dx, 2.0) - 0.5 * pow(dy, 2.0)) / side[i];
        H[3 * 4 + i] = -dy / side[i];
        H[4 * 4 + i] = (0.25 * pow(dy, 2.0) - 0.5 * pow(dx, 2.0)) / side[i];
    }
    memset(Kep, 0, 144 * sizeof(double));
    double root = sqrt(1.0 / 3.0);
    double J[4] = {0, 0, 0, 0};
    int lu_flag = 0, piv[2] = {0, 1}; /* J is declared once: fs.cpp:633 */
    for (int ii = 0; ii < 2; ii++) {
        double r = pow(-1.0, ii) * root;
        for (int jj = 0; jj < 2; jj++) {
            double s = pow(-1.0, jj) * root;
            J[0] = (dphi[0] + dphi[4]) * s - dphi[0] + dphi[4];
            J[1] = (dphi[1] + dphi[5]) * s - dphi[1] + dphi[5];
            J[2] = (dphi[0] + dphi[4]) * r - dphi[2] + dphi[6];
            J[3] = (dphi[1] + dphi[5]) * r - dphi[3] + dphi[7];
            for (int i = 0; i < 4; i++) J[i] *= 0.25;
            double det;
            if (quirks & FSO_QUIRK_DET_LU)
                det = det2_libmesh(J, &lu_flag, piv); /* flag survives across GPs */
            else
                det = J[0] * J[3] - J[1] * J[2];
            double Jinv[4] = {J[3], -J[1], -J[2], J[0]};
            double di = 1.0 / det;
            for (int i = 0; i < 4; i++) Jinv[i] *= di;
            double B[36], BtD[36], tmp[144];
            eval_b_quad(H, r, s, Jinv, B);
            mtm(3, 12, 3, B, Dp, BtD);   /* fs.cpp:680 */
            mm(12, 3, 12, BtD, B, tmp);  /* fs.cpp:681 */
            for (int i = 0; i < 144; i++) Kep[i] += tmp[i] * det;
        }
    }
}

/* ------------------------------------------------------------------ */
/* fs.cpp:999-1053 constructStiffnessMatrix (node-major 6x6 blocks)    */
/* ------------------------------------------------------------------ */
static void construct_shell(int nen, const double *Kem, const double *Kep, double *K)
{
    int n6 = 6 * nen, nm = 2 * nen, np = 3 * nen;
    memset(K, 0, sizeof(double) * n6 * n6);
    for (int i = 0; i < nen; i++)
        for (int j = 0; j < nen; j++) {
            for (int a = 0; a < 2; a++)
                for (int b = 0; b < 2; b++)
                    K[(6 * i + a) * n6 + 6 * j + b] = Kem[(2 * i + a) * nm + 2 * j + b];
            for (int a = 0; a < 3; a++)
                for (int b = 0; b < 3; b++)
                    K[(6 * i + 2 + a) * n6 + 6 * j + 2 + b] = Kep[(3 * i + a) * np + 3 * j + b];
            /* fs.cpp:1036-1051 drilling term on EVERY block, max/1000 */
            double mx = Kem[(2 * i) * nm + 2 * j];
            mx = fmax(mx, Kem[(2 * i + 1) * nm + 2 * j + 1]);
            mx = fmax(mx, Kep[(3 * i) * np + 3 * j]);
            mx = fmax(mx, Kep[(3 * i + 1) * np + 3 * j + 1]);
            mx = fmax(mx, Kep[(3 * i + 2) * np + 3 * j + 2]);
            K[(6 * i + 5) * n6 + 6 * j + 5] = mx / 1000.0;
        }
}

/* ------------------------------------------------------------------ */
/* fs.cpp:1061-1102 localToGlobalTrafo, rotation part (node-major)     */
/* ------------------------------------------------------------------ */
static void local_to_global(int nen, const double *trafo, double *K)
{
    int n6 = 6 * nen;
    double T[36];
    memset(T, 0, sizeof T);
    for (int k = 0; k < 2; k++)
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++)
                T[(3 * k + i) * 6 + 3 * k + j] = trafo[i * 3 + j];
    for (int i = 0; i < nen; i++)
        for (int j = 0; j < nen; j++) {
            double S[36], ST[36], R[36];
            for (int k = 0; k < 6; k++)
                for (int l = 0; l < 6; l++)
                    S[k * 6 + l] = K[(6 * i + k) * n6 + 6 * j + l];
            mm(6, 6, 6, S, T, ST);   /* fs.cpp:1094 */
            mtm(6, 6, 6, T, ST, R);  /* fs.cpp:1095 */
            for (int k = 0; k < 6; k++)
                for (int l = 0; l < 6; l++)
                    K[(6 * i + k) * n6 + 6 * j + l] = R[k * 6 + l];
        }
}

/*
 * One element, full pipeline fs.cpp:1211-1221.
 *   layout = 0 : node-major   K[(6i+a)(6n) + 6j+b]      (before fs.cpp:1105)
 *   layout = 1 : variable-major K[(n a + i)(6n) + n b + j] (after fs.cpp:1105-1109)
 * Returns nen.
 */
int fso_element_stiffness(int type, const double *xyz, double nu, double em, double thickness,
                          int quirks, int layout, double *Kout)
{
    double Dm[9], Dp[9];
    fso_material(nu, em, thickness, Dm, Dp);
    int nen = (type == FSO_TRI3) ? 3 : 4;
    double trafo[9], loc[12], dphi[8], area;
    double Kem[64], Kep[144], K[576];
    init_element(type, xyz, trafo, loc, dphi, &area);
    calc_plane(type, loc, dphi, area, Dm, thickness, quirks, Kem);
    calc_plate(type, dphi, area, Dp, quirks, Kep);
    construct_shell(nen, Kem, Kep, K);
    local_to_global(nen, trafo, K);
    int n6 = 6 * nen;
    if (layout == 0) {
        memcpy(Kout, K, sizeof(double) * n6 * n6);
    } else {
        for (int al = 0; al < 6; al++)
            for (int be = 0; be < 6; be++)
                for (int i = 0; i < nen; i++)
                    for (int j = 0; j < nen; j++)
                        Kout[(nen * al + i) * n6 + nen * be + j] = K[(6 * i + al) * n6 + 6 * j + be];
    }
    return nen;
}

/* geometry only, for kernel-level checks: trafo(9) loc(12) dphi(8) area(1) */
void fso_element_geometry(int type, const double *xyz, double *trafo, double *loc, double *dphi,
                          double *area)
{
    memset(loc, 0, 12 * sizeof(double));
    memset(dphi, 0, 8 * sizeof(double));
    init_element(type, xyz, trafo, loc, dphi, area);
}

/* local sub-matrices before superposition, for kernel-level checks */
void fso_element_parts(int type, const double *xyz, double nu, double em, double thickness,
                       int quirks, double *Kem, double *Kep)
{
    double Dm[9], Dp[9];
    fso_material(nu, em, thickness, Dm, Dp);
    double trafo[9], loc[12], dphi[8], area;
    init_element(type, xyz, trafo, loc, dphi, &area);
    calc_plane(type, loc, dphi, area, Dm, thickness, quirks, Kem);
    calc_plate(type, dphi, area, Dp, quirks, Kep);
}

/* ------------------------------------------------------------------ */
/* Stress resultants at the element centroid. */
/* NOT in the reference code: the thesis only states the formulas,     */
/*   membrane  sigma = Dm B u      doc/shellelements.tex:524           */
/*   bending   M = Dp B(x,y) w     doc/shellelements.tex:1394-1403     */
/* They are evaluated with the element's own strain-displacement       */
/* operators, i.e. exactly the B of calc_plane / calc_plate above      */
/* (Tri-3: Y * B~ of fs.cpp:578-595 at L1=L2=1/3; Quad-4: Bm*G of      */
/* fs.cpp:490-529 and evalBQuad at xi=eta=0, plain 2x2 determinant),   */
/* on the nodal displacements rotated into the element frame with the  */
/* transformation of fs.cpp:378-384 (u_loc = T u_glob, the inverse of  */
/* fs.cpp:1094-1095).  sols[6*node+var] as printed by fs.cpp:163-169.  */
/* out[6e..6e+5] = sigma_xx, sigma_yy, sigma_xy, M_x, M_y, M_xy in the */
/* local axes of element e.                                            */
/* ------------------------------------------------------------------ */
void fso_recover_resultants(const double *xyz, int64_t n_elem, const int32_t *etype,
                            const int64_t *eptr, const int32_t *enodes, double nu, double em,
                            double thickness, int quirks, const double *sols, double *out)
{
    double Dm[9], Dp[9];
    fso_material(nu, em, thickness, Dm, Dp);
    for (int64_t e = 0; e < n_elem; e++) {
        int type = etype[e];
        int nen = (int)(eptr[e + 1] - eptr[e]);
        const int32_t *en = enodes + eptr[e];
        double X[12], trafo[9], loc[12], dphi[8], area;
        for (int i = 0; i < nen; i++)
            for (int d = 0; d < 3; d++) X[3 * i + d] = xyz[3 * (int64_t)en[i] + d];
        init_element(type, X, trafo, loc, dphi, &area);
        double um[8], wp[12]; /* membrane (u,v) and plate (w,tx,ty) unknowns, node-major */
        for (int i = 0; i < nen; i++) {
            const double *g = sols + 6 * (int64_t)en[i];
            double ul[3], tl[3];
            mm(3, 3, 1, trafo, g, ul);
            mm(3, 3, 1, trafo, g + 3, tl);
            um[2 * i + 0] = ul[0];
            um[2 * i + 1] = ul[1];
            wp[3 * i + 0] = ul[2];
            wp[3 * i + 1] = tl[0];
            wp[3 * i + 2] = tl[1];
        }
        double eps[3], kap[3];
        if (type == FSO_TRI3) {
            double x12 = dphi[0], y12 = dphi[1], x31 = dphi[2], y31 = dphi[3], x23 = dphi[4],
                   y23 = dphi[5];
            double B[18];
            memset(B, 0, sizeof B);
            B[0 * 6 + 0] = y23;  B[0 * 6 + 2] = y31;  B[0 * 6 + 4] = y12;
            B[1 * 6 + 1] = -x23; B[1 * 6 + 3] = -x31; B[1 * 6 + 5] = -x12;
            B[2 * 6 + 0] = -x23; B[2 * 6 + 1] = y23;
            B[2 * 6 + 2] = -x31; B[2 * 6 + 3] = y31;
            B[2 * 6 + 4] = -x12; B[2 * 6 + 5] = y12;
            double s = 1.0 / (2.0 * area);
            for (int i = 0; i < 18; i++) B[i] *= s;
            mm(3, 6, 1, B, um, eps);
            double side[3], Bt[27], Y[9], kt[3];
            for (int i = 0; i < 3; i++)
                side[i] = pow(dphi[2 * i], 2.0) + pow(dphi[2 * i + 1], 2.0);
            eval_b_tri(side, 1.0 / 3.0, 1.0 / 3.0, dphi, Bt);
            Y[0] = pow(y23, 2.0); Y[1] = pow(y31, 2.0); Y[2] = y23 * y31;
            Y[3] = pow(x23, 2.0); Y[4] = pow(x31, 2.0); Y[5] = x31 * x23;
            Y[6] = -2.0 * x23 * y23;
            Y[7] = (quirks & FSO_QUIRK_Y21) ? -2.0 * x31 * x31 : -2.0 * x31 * y31;
            Y[8] = -x23 * y31 - x31 * y23;
            double sc = 1.0 / (4.0 * pow(area, 2.0));
            for (int i = 0; i < 9; i++) Y[i] *= sc;
            mm(3, 9, 1, Bt, wp, kt);
            mm(3, 3, 1, Y, kt, kap);
        } else {
            double dhdr[4] = {-0.25, 0.25, 0.25, -0.25}, dhds[4] = {-0.25, -0.25, 0.25, 0.25};
            double J[4] = {0, 0, 0, 0};
            for (int i = 0; i < 4; i++) {
                J[0] += dhdr[i] * loc[0 * 4 + i];
                J[1] += dhdr[i] * loc[1 * 4 + i];
                J[2] += dhds[i] * loc[0 * 4 + i];
                J[3] += dhds[i] * loc[1 * 4 + i];
            }
            double di = 1.0 / (J[0] * J[3] - J[1] * J[2]);
            double Bm[12], G[32], B[24];
            memset(Bm, 0, sizeof Bm);
            memset(G, 0, sizeof G);
            Bm[0 * 4 + 0] = J[3];  Bm[0 * 4 + 1] = -J[1];
            Bm[1 * 4 + 2] = -J[2]; Bm[1 * 4 + 3] = J[0];
            Bm[2 * 4 + 0] = -J[2]; Bm[2 * 4 + 1] = J[0];
            Bm[2 * 4 + 2] = J[3];  Bm[2 * 4 + 3] = -J[1];
            for (int i = 0; i < 12; i++) Bm[i] *= di;
            for (int i = 0; i < 4; i++) {
                G[0 * 8 + 2 * i] = dhdr[i];
                G[1 * 8 + 2 * i] = dhds[i];
                G[2 * 8 + 1 + 2 * i] = dhdr[i];
                G[3 * 8 + 1 + 2 * i] = dhds[i];
            }
            mm(3, 4, 8, Bm, G, B);
            mm(3, 8, 1, B, um, eps);
            double H[20], Jp[4], Jinv[4], Bp[36];
            for (int i = 0; i < 4; i++) {
                double dx = dphi[2 * i], dy = dphi[2 * i + 1];
                double sd = pow(dx, 2.0) + pow(dy, 2.0);
                H[0 * 4 + i] = -dx / sd;
                H[1 * 4 + i] = 0.75 * dx * dy / sd;
                H[2 * 4 + i] = (0.25 * pow(dx, 2.0) - 0.5 * pow(dy, 2.0)) / sd;
                H[3 * 4 + i] = -dy / sd;
                H[4 * 4 + i] = (0.25 * pow(dy, 2.0) - 0.5 * pow(dx, 2.0)) / sd;
            }
            Jp[0] = 0.25 * (-dphi[0] + dphi[4]);
            Jp[1] = 0.25 * (-dphi[1] + dphi[5]);
            Jp[2] = 0.25 * (-dphi[2] + dphi[6]);
            Jp[3] = 0.25 * (-dphi[3] + dphi[7]);
            double dp = 1.0 / (Jp[0] * Jp[3] - Jp[1] * Jp[2]);
            Jinv[0] = Jp[3] * dp; Jinv[1] = -Jp[1] * dp; Jinv[2] = -Jp[2] * dp; Jinv[3] = Jp[0] * dp;
            eval_b_quad(H, 0.0, 0.0, Jinv, Bp);
            mm(3, 12, 1, Bp, wp, kap);
        }
        mm(3, 3, 1, Dm, eps, out + 6 * e);
        mm(3, 3, 1, Dp, kap, out + 6 * e + 3);
    }
}

/* ------------------------------------------------------------------ */
/* DOF numbering (libMesh DofMap, var-major distribution, one variable */
/* group of six FIRST/LAGRANGE variables): node bases are handed out   */
/* in first-encounter order over elements in id order, nodes in local  */
/* order; the six DOFs of a node are contiguous (u,v,w,tx,ty,tz).      */
/*   mode 0 = FIRST_ENCOUNTER (libMesh), 1 = NODE_ID                   */
/* dofnode[n] = position of node n (>=0) or -1 for a node no element   */
/* references.  Returns number of numbered nodes.                      */
/* ------------------------------------------------------------------ */
int64_t fso_dof_order(int64_t n_nodes, int64_t n_elem, const int64_t *eptr, const int32_t *enodes,
                      int mode, int32_t *dofnode)
{
    for (int64_t i = 0; i < n_nodes; i++) dofnode[i] = -1;
    int64_t next = 0;
    if (mode == 0) {
        for (int64_t e = 0; e < n_elem; e++)
            for (int64_t k = eptr[e]; k < eptr[e + 1]; k++)
                if (dofnode[enodes[k]] < 0) dofnode[enodes[k]] = (int32_t)next++;
    } else {
        for (int64_t e = 0; e < n_elem; e++)
            for (int64_t k = eptr[e]; k < eptr[e + 1]; k++) dofnode[enodes[k]] = 0;
        for (int64_t i = 0; i < n_nodes; i++)
            if (dofnode[i] == 0) dofnode[i] = (int32_t)next++;
    }
    return next;
}

/* ------------------------------------------------------------------ */
/* Dirichlet sets, fs.cpp:90-120: side boundary ids {0,20} fix u,v,w;  */
/* {1,21} fix all six.  bc rows = (element, side, id); side s joins    */
/* local nodes s and (s+1)%nen (src/meshgen/main_all.cpp:276-281).     */
/* mask[n] bit v set = variable v of node n constrained.               */
/* ------------------------------------------------------------------ */
void fso_constraint_mask(int64_t n_nodes, const int64_t *eptr, const int32_t *enodes, int64_t n_bc,
                         const int32_t *bc, uint8_t *mask)
{
    memset(mask, 0, (size_t)n_nodes);
    for (int64_t i = 0; i < n_bc; i++) {
        int32_t e = bc[3 * i], s = bc[3 * i + 1], id = bc[3 * i + 2];
        uint8_t m = 0;
        if (id == 0 || id == 20) m = 0x07;
        if (id == 1 || id == 21) m = 0x3f;
        if (!m) continue;
        int nen = (int)(eptr[e + 1] - eptr[e]);
        mask[enodes[eptr[e] + s]] |= m;
        mask[enodes[eptr[e] + (s + 1) % nen]] |= m;
    }
}

/* ------------------------------------------------------------------ */
/* Sparsity: a dense 6x6 block for every ordered node pair sharing an  */
/* element (what libMesh's SparsityPattern builds for six fully        */
/* coupled nodal variables and what MatSetValues of the dense element  */
/* block, zeros included, fills: fs.cpp:1230).  Returned per NODE:     */
/* nptr (n_dofnodes+1), nadj (sorted dof-node ids).  Scalar CSR:       */
/* row 6p+a has columns 6q+b for q in nadj[p], b = 0..5.               */
/* Two-call protocol: nadj == NULL -> only nptr is filled.             */
/* ------------------------------------------------------------------ */
static int cmp_i32(const void *a, const void *b)
{
    int32_t x = *(const int32_t *)a, y = *(const int32_t *)b;
    return (x > y) - (x < y);
}

int64_t fso_node_pattern(int64_t n_dofnodes, int64_t n_elem, const int64_t *eptr,
                         const int32_t *enodes, const int32_t *dofnode, int64_t *nptr,
                         int32_t *nadj)
{
    int64_t *cnt = (int64_t *)calloc((size_t)n_dofnodes + 1, sizeof(int64_t));
    for (int64_t e = 0; e < n_elem; e++) {
        int nen = (int)(eptr[e + 1] - eptr[e]);
        for (int64_t k = eptr[e]; k < eptr[e + 1]; k++) cnt[dofnode[enodes[k]] + 1] += nen;
    }
    for (int64_t i = 0; i < n_dofnodes; i++) cnt[i + 1] += cnt[i];
    int64_t tot = cnt[n_dofnodes];
    int32_t *cand = (int32_t *)malloc(sizeof(int32_t) * (size_t)(tot ? tot : 1));
    int64_t *fill = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n_dofnodes + 1));
    memcpy(fill, cnt, sizeof(int64_t) * (size_t)(n_dofnodes + 1));
    for (int64_t e = 0; e < n_elem; e++)
        for (int64_t k = eptr[e]; k < eptr[e + 1]; k++) {
            int32_t p = dofnode[enodes[k]];
            for (int64_t l = eptr[e]; l < eptr[e + 1]; l++) cand[fill[p]++] = dofnode[enodes[l]];
        }
    int64_t nnzb = 0;
    nptr[0] = 0;
    for (int64_t p = 0; p < n_dofnodes; p++) {
        int64_t b = cnt[p], n = cnt[p + 1] - b;
        qsort(cand + b, (size_t)n, sizeof(int32_t), cmp_i32);
        int64_t u = 0;
        for (int64_t i = 0; i < n; i++)
            if (i == 0 || cand[b + i] != cand[b + i - 1]) {
                if (nadj) nadj[nnzb + u] = cand[b + i];
                u++;
            }
        nnzb += u;
        nptr[p + 1] = nnzb;
    }
    free(cnt); free(cand); free(fill);
    return nnzb;
}

/* expand the node pattern to scalar CSR (rowptr int64, colidx int32) */
void fso_expand_csr(int64_t n_dofnodes, const int64_t *nptr, const int32_t *nadj, int64_t *rowptr,
                    int32_t *colidx)
{
    for (int64_t p = 0; p < n_dofnodes; p++) {
        int64_t deg = nptr[p + 1] - nptr[p];
        for (int a = 0; a < 6; a++) {
            int64_t r = 6 * p + a;
            int64_t base = 36 * nptr[p] + a * 6 * deg;
            rowptr[r] = base;
            if (colidx)
                for (int64_t j = 0; j < deg; j++)
                    for (int b = 0; b < 6; b++)
                        colidx[base + 6 * j + b] = 6 * nadj[nptr[p] + j] + b;
        }
    }
    rowptr[6 * n_dofnodes] = 36 * nptr[n_dofnodes];
}

/* ------------------------------------------------------------------ */
/* fs.cpp:1160-1233 assemble_elasticity: element loop in id order,     */
/* Dirichlet element constraint (libMesh                               */
/* constrain_element_matrix_and_vector for homogeneous Dirichlet rows: */
/* zero constrained rows and columns, put 1 on their diagonal, zero    */
/* their rhs), then ADD into the global matrix / rhs.                  */
/* vals are in scalar-CSR order as produced by fso_expand_csr.         */
/* forces: n_nodes x 6 by mesh node id.  rhs: 6*n_dofnodes.            */
/* threads <= 1: strictly serial, element-id order (the parity path).  */
/* threads  > 1: OpenMP over elements with atomic adds (timing only).  */
/* ------------------------------------------------------------------ */
static int64_t find_adj(const int32_t *adj, int64_t n, int32_t key)
{
    int64_t lo = 0, hi = n - 1;
    while (lo <= hi) {
        int64_t mid = (lo + hi) / 2;
        if (adj[mid] < key) lo = mid + 1;
        else if (adj[mid] > key) hi = mid - 1;
        else return mid;
    }
    return -1;
}

void fso_assemble(int64_t n_nodes, const double *xyz, int64_t n_elem, const int32_t *etype,
                  const int64_t *eptr, const int32_t *enodes, const int32_t *dofnode,
                  const uint8_t *mask, const double *forces, double nu, double em,
                  double thickness, int quirks, int64_t n_dofnodes, const int64_t *nptr,
                  const int32_t *nadj, double *vals, double *rhs, int threads)
{
    (void)n_nodes;
    memset(vals, 0, sizeof(double) * (size_t)(36 * nptr[n_dofnodes]));
    memset(rhs, 0, sizeof(double) * (size_t)(6 * n_dofnodes));
    /* fs.cpp:1193 processedNodes: every node loads the rhs exactly once */
    uint8_t *processed = (uint8_t *)calloc((size_t)n_nodes, 1);
    int par = threads > 1;
#ifdef _OPENMP
    if (par) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(static) if (par)
    for (int64_t e = 0; e < n_elem; e++) {
        int type = etype[e];
        int nen = (int)(eptr[e + 1] - eptr[e]);
        int n6 = 6 * nen;
        const int32_t *en = enodes + eptr[e];
        double X[12], Ke[576], Fe[24];
        for (int i = 0; i < nen; i++)
            for (int d = 0; d < 3; d++) X[3 * i + d] = xyz[3 * (int64_t)en[i] + d];
        /* variable-major, as handed to libMesh: fs.cpp:1105-1109 */
        fso_element_stiffness(type, X, nu, em, thickness, quirks, 1, Ke);
        /* fs.cpp:1118-1153 contribRHS */
        memset(Fe, 0, sizeof Fe);
        for (int s = 0; s < nen; s++) {
            int first;
            if (par) {
                uint8_t old;
#pragma omp atomic capture
                { old = processed[en[s]]; processed[en[s]] = 1; }
                first = !old;
            } else {
                first = !processed[en[s]];
                processed[en[s]] = 1;
            }
            if (first && forces)
                for (int i = 0; i < 6; i++) Fe[s + nen * i] = forces[6 * (int64_t)en[s] + i];
        }
        /* fs.cpp:1227 constrain_element_matrix_and_vector; local dof (var al, node i) = nen*al+i */
        for (int al = 0; al < 6; al++)
            for (int i = 0; i < nen; i++)
                if (mask[en[i]] & (1u << al)) {
                    int c = nen * al + i;
                    for (int k = 0; k < n6; k++) {
                        Ke[c * n6 + k] = 0.0;
                        Ke[k * n6 + c] = 0.0;
                    }
                    Ke[c * n6 + c] = 1.0;
                    Fe[c] = 0.0;
                }
        /* fs.cpp:1230-1231 add_matrix / add_vector; global dof = 6*dofnode + var (a12) */
        for (int i = 0; i < nen; i++) {
            int32_t p = dofnode[en[i]];
            int64_t deg = nptr[p + 1] - nptr[p];
            for (int j = 0; j < nen; j++) {
                int32_t q = dofnode[en[j]];
                int64_t pos = find_adj(nadj + nptr[p], deg, q);
                for (int al = 0; al < 6; al++)
                    for (int be = 0; be < 6; be++) {
                        double v = Ke[(nen * al + i) * n6 + nen * be + j];
                        double *dst = &vals[36 * nptr[p] + al * 6 * deg + 6 * pos + be];
                        if (par) {
#pragma omp atomic
                            *dst += v;
                        } else
                            *dst += v;
                    }
            }
            for (int al = 0; al < 6; al++) {
                double *dst = &rhs[6 * (int64_t)p + al];
                if (par) {
#pragma omp atomic
                    *dst += Fe[nen * al + i];
                } else
                    *dst += Fe[nen * al + i];
            }
        }
    }
    free(processed);
}

/* ------------------------------------------------------------------ */
/* Krylov solve standing in for KSPSolve (fs.cpp:138) with             */
/* -ksp_type cg -pc_type jacobi | pbjacobi: textbook preconditioned    */
/* CG on the 6x6-block CSR above.                                      */
/*   pc: 0 none, 1 Jacobi (point diagonal), 2 block-6 Jacobi           */
/*   norm_type: 0 = ||r||_2 / ||b||_2, 1 = ||M^-1 r||_2 / ||M^-1 b||_2 */
/*   x holds the initial guess on entry (warm start, libMesh default). */
/* Returns iterations done; *relres = final relative residual; <0 on   */
/* breakdown (p.Ap <= 0).                                              */
/* ------------------------------------------------------------------ */
static void spmv(int64_t nn, const int64_t *nptr, const int32_t *nadj, const double *vals,
                 const double *x, double *y)
{
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < nn; p++) {
        int64_t deg = nptr[p + 1] - nptr[p];
        const int32_t *adj = nadj + nptr[p];
        for (int a = 0; a < 6; a++) {
            const double *row = vals + 36 * nptr[p] + a * 6 * deg;
            double s = 0.0;
            for (int64_t j = 0; j < deg; j++) {
                const double *xx = x + 6 * (int64_t)adj[j];
                for (int b = 0; b < 6; b++) s += row[6 * j + b] * xx[b];
            }
            y[6 * p + a] = s;
        }
    }
}

static double dotp(int64_t n, const double *a, const double *b)
{
    double s = 0.0;
#pragma omp parallel for reduction(+ : s) schedule(static)
    for (int64_t i = 0; i < n; i++) s += a[i] * b[i];
    return s;
}

/* in-place inverse of a 6x6 by Gauss-Jordan with partial pivoting */
static int inv6(double *A)
{
    double M[6][12];
    for (int i = 0; i < 6; i++)
        for (int j = 0; j < 6; j++) {
            M[i][j] = A[i * 6 + j];
            M[i][6 + j] = (i == j);
        }
    for (int c = 0; c < 6; c++) {
        int piv = c;
        for (int r = c + 1; r < 6; r++)
            if (fabs(M[r][c]) > fabs(M[piv][c])) piv = r;
        if (M[piv][c] == 0.0) return -1;
        if (piv != c)
            for (int j = 0; j < 12; j++) {
                double t = M[c][j]; M[c][j] = M[piv][j]; M[piv][j] = t;
            }
        double d = 1.0 / M[c][c];
        for (int j = 0; j < 12; j++) M[c][j] *= d;
        for (int r = 0; r < 6; r++)
            if (r != c) {
                double f = M[r][c];
                if (f != 0.0)
                    for (int j = 0; j < 12; j++) M[r][j] -= f * M[c][j];
            }
    }
    for (int i = 0; i < 6; i++)
        for (int j = 0; j < 6; j++) A[i * 6 + j] = M[i][6 + j];
    return 0;
}

static void apply_pc(int pc, int64_t nn, const double *minv, const double *r, double *z)
{
    if (pc == 0) {
        memcpy(z, r, sizeof(double) * (size_t)(6 * nn));
    } else if (pc == 1) {
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < 6 * nn; i++) z[i] = minv[i] * r[i];
    } else {
#pragma omp parallel for schedule(static)
        for (int64_t p = 0; p < nn; p++)
            for (int a = 0; a < 6; a++) {
                double s = 0.0;
                for (int b = 0; b < 6; b++) s += minv[36 * p + 6 * a + b] * r[6 * p + b];
                z[6 * p + a] = s;
            }
    }
}

int64_t fso_pcg(int64_t n_dofnodes, const int64_t *nptr, const int32_t *nadj, const double *vals,
                const double *b, double *x, int pc, int norm_type, double rtol, int64_t max_its,
                int threads, double *relres)
{
    int64_t nn = n_dofnodes, n = 6 * nn;
#ifdef _OPENMP
    omp_set_num_threads(threads > 0 ? threads : 1);
#endif
    double *r = (double *)malloc(sizeof(double) * (size_t)n);
    double *z = (double *)malloc(sizeof(double) * (size_t)n);
    double *p = (double *)malloc(sizeof(double) * (size_t)n);
    double *q = (double *)malloc(sizeof(double) * (size_t)n);
    double *minv = NULL;
    if (pc == 1) {
        minv = (double *)malloc(sizeof(double) * (size_t)n);
        for (int64_t pn = 0; pn < nn; pn++) {
            int64_t deg = nptr[pn + 1] - nptr[pn];
            int64_t pos = find_adj(nadj + nptr[pn], deg, (int32_t)pn);
            for (int a = 0; a < 6; a++)
                minv[6 * pn + a] = 1.0 / vals[36 * nptr[pn] + a * 6 * deg + 6 * pos + a];
        }
    } else if (pc == 2) {
        minv = (double *)malloc(sizeof(double) * (size_t)(36 * nn));
        for (int64_t pn = 0; pn < nn; pn++) {
            int64_t deg = nptr[pn + 1] - nptr[pn];
            int64_t pos = find_adj(nadj + nptr[pn], deg, (int32_t)pn);
            for (int a = 0; a < 6; a++)
                for (int c = 0; c < 6; c++)
                    minv[36 * pn + 6 * a + c] = vals[36 * nptr[pn] + a * 6 * deg + 6 * pos + c];
            inv6(minv + 36 * pn);
        }
    }
    int64_t it = 0;
    int64_t status = 0;
    /* reference norm: b (or M^-1 b), PETSc's default for a non-zero initial guess as well */
    double bn;
    if (norm_type == 1) {
        apply_pc(pc, nn, minv, b, z);
        bn = sqrt(dotp(n, z, z));
    } else
        bn = sqrt(dotp(n, b, b));
    spmv(nn, nptr, nadj, vals, x, q);
    for (int64_t i = 0; i < n; i++) r[i] = b[i] - q[i];
    apply_pc(pc, nn, minv, r, z);
    double rz = dotp(n, r, z);
    double rn = (norm_type == 1) ? sqrt(dotp(n, z, z)) : sqrt(dotp(n, r, r));
    memcpy(p, z, sizeof(double) * (size_t)n);
    if (bn == 0.0) {
        memset(x, 0, sizeof(double) * (size_t)n);
        rn = 0.0;
        bn = 1.0;
    }
    while (rn > rtol * bn && it < max_its) {
        spmv(nn, nptr, nadj, vals, p, q);
        double pq = dotp(n, p, q);
        if (!(pq > 0.0)) { status = -1; break; }
        double alpha = rz / pq;
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; i++) {
            x[i] += alpha * p[i];
            r[i] -= alpha * q[i];
        }
        apply_pc(pc, nn, minv, r, z);
        double rz_new = dotp(n, r, z);
        rn = (norm_type == 1) ? sqrt(dotp(n, z, z)) : sqrt(dotp(n, r, r));
        double beta = rz_new / rz;
        rz = rz_new;
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; i++) p[i] = z[i] + beta * p[i];
        it++;
    }
    if (relres) *relres = rn / bn;
    free(r); free(z); free(p); free(q); free(minv);
    return status < 0 ? status : it;
}

/* y = A x on the oracle's CSR, exported for SpMV parity checks */
void fso_spmv(int64_t n_dofnodes, const int64_t *nptr, const int32_t *nadj, const double *vals,
              const double *x, double *y, int threads)
{
#ifdef _OPENMP
    omp_set_num_threads(threads > 0 ? threads : 1);
#endif
    spmv(n_dofnodes, nptr, nadj, vals, x, y);
}

/* fs.cpp:140-141,163-169: sols[6*node_id + var] from the DOF-ordered solution */
void fso_gather_solution(int64_t n_nodes, const int32_t *dofnode, const double *x, double *sols)
{
    for (int64_t i = 0; i < n_nodes; i++)
        for (int v = 0; v < 6; v++)
            sols[6 * i + v] = dofnode[i] >= 0 ? x[6 * (int64_t)dofnode[i] + v] : 0.0;
}

int fso_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// fs_elements.cuh -- FP64 device routines for the flat-shell element stiffness.
//
// One thread owns ONE NODE ROW of one element: it forms the membrane (2 x 2n) and plate (3 x 3n)
// row slices of the local matrices in registers, superposes them with the drilling term into the
// n 6x6 node blocks of that row, rotates each block to global axes and hands it to a sink
// (coloured scatter, row gather or a dense debug buffer).  Nothing is kept in local memory: all
// loops are fully unrolled over compile-time node indices.
//
// Reference functions restated here (precice/fem-shell, src/fem-shell/fem-shell.cpp = "fs.cpp"):
//   initElement               fs.cpp:306-433      tri_geom / quad_geom
//   calcPlane                 fs.cpp:443-542      *_membrane_row
//   calcPlate                 fs.cpp:551-688      *_plate_row
//   evalBTri                  fs.cpp:698-891      tri_gp_terms / tri_bcols
//   evalBQuad                 fs.cpp:901-990      quad_bcols
//   constructStiffnessMatrix  fs.cpp:999-1053     emit_row_blocks (superposition + drilling)
//   localToGlobalTrafo        fs.cpp:1061-1102    rotate_block
// including the reference's arithmetic quirks: Y(2,1) at fs.cpp:586 and the
// in-place LU that libMesh's DenseMatrix::det() leaves behind at fs.cpp:512 and fs.cpp:652.
#pragma once
#include <cuda_runtime.h>

#include <cmath>

namespace fs {

struct ElemConst {
    double dm11, dm12, dm33;  // Dm = E/(1-nu^2) * [[1,nu,0],[nu,1,0],[0,0,(1-nu)/2]]   fs.cpp:289
    double dp11, dp12, dp33;  // Dp = E t^3/(12(1-nu^2)) * (same)                        fs.cpp:293
    double thickness;
    int quirks;               // FS_QUIRK_* bits
};

// Gauss-point loops of the run-time-row quad kernels: 4 = fully unrolled (r, s fold into the shape-function constants),
// 1 = rolled (a quarter of the code: the values pass is short of instruction-cache hits, profiles/r02k_ncu_full.txt)
#ifndef FS_QUAD_GP_UNROLL
#define FS_QUAD_GP_UNROLL 4
#endif
constexpr int QUAD_GP_UNROLL = FS_QUAD_GP_UNROLL;

__constant__ ElemConst c_el;   // one copy per translation unit that includes this header (no -rdc)

// fs.cpp:273-294 initMaterialMatrices
static inline ElemConst make_elem_const(double nu, double E, double thickness, int quirks)
{
    ElemConst h;
    const double fm = E / (1.0 - nu * nu);
    const double fp = E * pow(thickness, 3.0) / (12.0 * (1.0 - nu * nu));
    h.dm11 = 1.0 * fm; h.dm12 = nu * fm; h.dm33 = ((1.0 - nu) / 2.0) * fm;
    h.dp11 = 1.0 * fp; h.dp12 = nu * fp; h.dp33 = ((1.0 - nu) / 2.0) * fp;
    h.thickness = thickness;
    h.quirks = quirks;
    return h;
}

// uploads into THIS translation unit's c_el (static: one instance per including .cu file)
static inline cudaError_t upload_elem_const_tu(const ElemConst &h, cudaStream_t st)
{
    return cudaMemcpyToSymbolAsync(c_el, &h, sizeof h, 0, cudaMemcpyHostToDevice, st);
}

#define FS_Q_Y21 1
#define FS_Q_DETLU 2

// ---------------------------------------------------------------------------------------------
// libMesh DenseMatrix::det() on a 2x2 (see oracle/fs_oracle.c det2_libmesh for the long form)
// ---------------------------------------------------------------------------------------------
struct Lu2State {
    bool done;
    bool swapped;
};

__device__ __forceinline__ double det2_quirk(double &j00, double &j01, double &j10, double &j11,
                                             Lu2State &st, bool quirk)
{
    if (!quirk) return j00 * j11 - j01 * j10;
    if (!st.done) {
        st.swapped = fabs(j00) < fabs(j10);
        if (st.swapped) {
            double a = j00, b = j01;
            j00 = j10; j01 = j11;
            j10 = a;   j11 = b;
        }
        double dinv = 1.0 / j00;
        j01 *= dinv;
        j11 -= j10 * j01;
        st.done = true;
    }
    double d = j00 * j11;
    return st.swapped ? -d : d;
}

// ---------------------------------------------------------------------------------------------
// geometry, fs.cpp:306-433
// ---------------------------------------------------------------------------------------------
struct TriGeom {
    double T[3][3];  // rows = local x, y, z axes in global coordinates
    double x12, y12, x31, y31, x23, y23;
    double area;
};

struct QuadGeom {
    double T[3][3];
    double lx[4], ly[4];  // local node coordinates (not translated, fs.cpp:391)
    double dx[4], dy[4];  // dphi: p_k - p_{k+1}
    double area;
};

__device__ __forceinline__ void cross3(const double a[3], const double b[3], double c[3])
{
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = -a[0] * b[2] + a[2] * b[0];
    c[2] = a[0] * b[1] - a[1] * b[0];
}

__device__ __forceinline__ void unit3(double a[3])
{
    const double li = 1.0 / sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
    a[0] *= li; a[1] *= li; a[2] *= li;
}

// X: 3 nodes x 3 coordinates
__device__ __forceinline__ void tri_geom(const double X[9], TriGeom &g)
{
    double U[3], V[3], W[3], Ug[3], Vg[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        Ug[i] = U[i] = X[3 + i] - X[i];
        Vg[i] = V[i] = X[6 + i] - X[i];
    }
    cross3(U, V, W);
    g.area = 0.5 * sqrt(W[0] * W[0] + W[1] * W[1] + W[2] * W[2]);
    unit3(U);
    unit3(W);
    cross3(W, U, V);
#pragma unroll
    for (int i = 0; i < 3; i++) {
        g.T[0][i] = U[i];
        g.T[1][i] = V[i];
        g.T[2][i] = W[i];
    }
    // local coordinates of B and C (A is the origin): fs.cpp:391
    double x2 = U[0] * Ug[0] + U[1] * Ug[1] + U[2] * Ug[2];
    double y2 = V[0] * Ug[0] + V[1] * Ug[1] + V[2] * Ug[2];
    double x3 = U[0] * Vg[0] + U[1] * Vg[1] + U[2] * Vg[2];
    double y3 = V[0] * Vg[0] + V[1] * Vg[1] + V[2] * Vg[2];
    g.x12 = -x2;      g.y12 = -y2;       // fs.cpp:406,409
    g.x31 = x3;       g.y31 = y3;        // fs.cpp:407,410
    g.x23 = x2 - x3;  g.y23 = y2 - y3;   // fs.cpp:408,411
}

// X: 4 nodes x 3 coordinates
__device__ __forceinline__ void quad_geom(const double X[12], QuadGeom &g)
{
    double U[3], V[3], W[3], R[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        double a = X[i], b = X[3 + i], c = X[6 + i], d = X[9 + i];
        double nI = a + 0.5 * (b - a);
        double nJ = b + 0.5 * (c - b);
        double nK = c + 0.5 * (d - c);
        double nL = d + 0.5 * (a - d);
        U[i] = nJ - nL;
        R[i] = nK - nI;
    }
    unit3(U);
    cross3(U, R, W);
    unit3(W);
    cross3(W, U, V);
#pragma unroll
    for (int i = 0; i < 3; i++) {
        g.T[0][i] = U[i];
        g.T[1][i] = V[i];
        g.T[2][i] = W[i];
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
        g.lx[k] = U[0] * X[3 * k] + U[1] * X[3 * k + 1] + U[2] * X[3 * k + 2];
        g.ly[k] = V[0] * X[3 * k] + V[1] * X[3 * k + 1] + V[2] * X[3 * k + 2];
    }
    double a2 = 0.0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        int n = (k + 1) & 3;
        g.dx[k] = g.lx[k] - g.lx[n];
        g.dy[k] = g.ly[k] - g.ly[n];
        a2 += g.lx[k] * g.ly[n] - g.lx[n] * g.ly[k];
    }
    g.area = 0.5 * a2;
}

// ---------------------------------------------------------------------------------------------
// membrane rows.  For both elements the strain-displacement columns of node k are
//   [px_k 0; 0 py_k; py_k px_k], so K_Ij = f * B_I^T Dm B_j has the closed form below.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void membrane_accum(double f, double pxI, double pyI, double pxj,
                                               double pyj, double m[2][2])
{
    const double d11 = c_el.dm11, d12 = c_el.dm12, d33 = c_el.dm33;
    m[0][0] += f * (pxI * d11 * pxj + pyI * d33 * pyj);
    m[0][1] += f * (pxI * d12 * pyj + pyI * d33 * pxj);
    m[1][0] += f * (pyI * d12 * pxj + pxI * d33 * pyj);
    m[1][1] += f * (pyI * d11 * pyj + pxI * d33 * pxj);
}

// the same sums with the factors that do not depend on j hoisted: 8 multiplications per Gauss point, then
// 8 FMAs per column block (membrane_accum: 20 FP64 instructions per block)
struct MembraneRow {
    double a, b, c, d, e, f;
};
__device__ __forceinline__ MembraneRow membrane_row_coeffs(double f, double pxI, double pyI)
{
    const double fx = f * pxI, fy = f * pyI;
    MembraneRow r;
    r.a = fx * c_el.dm11; r.b = fy * c_el.dm33; r.c = fx * c_el.dm12;
    r.d = fy * c_el.dm12; r.e = fx * c_el.dm33; r.f = fy * c_el.dm11;
    return r;
}
__device__ __forceinline__ void membrane_accum_row(const MembraneRow &r, double pxj, double pyj, double m[2][2])
{
    // FMA chains INTO the accumulator (a sum in parentheses added afterwards costs an extra DADD per entry)
    m[0][0] = fma(r.b, pyj, fma(r.a, pxj, m[0][0]));
    m[0][1] = fma(r.b, pxj, fma(r.c, pyj, m[0][1]));
    m[1][0] = fma(r.e, pyj, fma(r.d, pxj, m[1][0]));
    m[1][1] = fma(r.e, pxj, fma(r.f, pyj, m[1][1]));
}

// CST, fs.cpp:445-468
template <int I>
__device__ __forceinline__ void tri_membrane_row(const TriGeom &g, double Km[3][2][2])
{
    const double s = 1.0 / (2.0 * g.area);
    const double px[3] = {g.y23 * s, g.y31 * s, g.y12 * s};
    const double py[3] = {-g.x23 * s, -g.x31 * s, -g.x12 * s};
    const double f = c_el.thickness * g.area;
#pragma unroll
    for (int j = 0; j < 3; j++) {
        Km[j][0][0] = Km[j][0][1] = Km[j][1][0] = Km[j][1][1] = 0.0;
        membrane_accum(f, px[I], py[I], px[j], py[j], Km[j]);
    }
}

// bilinear quad, 2x2 Gauss, fs.cpp:469-541
template <int I>
__device__ __forceinline__ void quad_membrane_row(const QuadGeom &g, double Km[4][2][2])
{
    const double root = 0.57735026918962584;  // sqrt(1.0/3.0), fs.cpp:472
    const bool quirk = (c_el.quirks & FS_Q_DETLU) != 0;
#pragma unroll
    for (int j = 0; j < 4; j++) Km[j][0][0] = Km[j][0][1] = Km[j][1][0] = Km[j][1][1] = 0.0;
#pragma unroll
    for (int gp = 0; gp < 4; gp++) {
        const double r = (gp & 2) ? -root : root;  // ii: fs.cpp:484
        const double s = (gp & 1) ? -root : root;  // jj: fs.cpp:487
        const double dr[4] = {-0.25 * (1 - s), 0.25 * (1 - s), 0.25 * (1 + s), -0.25 * (1 + s)};
        const double ds[4] = {-0.25 * (1 - r), -0.25 * (1 + r), 0.25 * (1 + r), 0.25 * (1 - r)};
        double j00 = 0, j01 = 0, j10 = 0, j11 = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            j00 += dr[k] * g.lx[k];
            j01 += dr[k] * g.ly[k];
            j10 += ds[k] * g.lx[k];
            j11 += ds[k] * g.ly[k];
        }
        Lu2State st = {false, false};  // fs.cpp:504: resize() clears the LU flag at every GP
        const double det = det2_quirk(j00, j01, j10, j11, st, quirk);
        const double di = 1.0 / det;
        const double b00 = j11 * di, b01 = -j01 * di, b12 = -j10 * di, b13 = j00 * di;  // fs.cpp:515-518
        double px[4], py[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            px[k] = b00 * dr[k] + b01 * ds[k];
            py[k] = b12 * dr[k] + b13 * ds[k];
        }
        const double f = det * c_el.thickness;
#pragma unroll
        for (int j = 0; j < 4; j++) membrane_accum(f, px[I], py[I], px[j], py[j], Km[j]);
    }
}

// run-time node row (callers keep I warp-uniform): same arithmetic, one copy of the code
// selp.f64 spelled out: nvcc turns nested ?: on doubles into divergent branch trees (BSSY/BRA/BSYNC), which cost
// the gather kernel 13 % of its issue slots in branch-resolving stalls (profiles/r01h_asm_notes.txt)
__device__ __forceinline__ double sel_f64(bool p, double a, double b)
{
    double d;
    asm("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %3, 0;\n\tselp.f64 %0, %1, %2, q;\n\t}" : "=d"(d) : "d"(a), "d"(b), "r"((int)p));
    return d;
}
__device__ __forceinline__ double pick3(int I, const double v[3]) { return sel_f64(I == 0, v[0], sel_f64(I == 1, v[1], v[2])); }
__device__ __forceinline__ double pick4(int I, const double v[4])
{
    return sel_f64(I < 2, sel_f64(I == 0, v[0], v[1]), sel_f64(I == 2, v[2], v[3]));
}

__device__ __forceinline__ void tri_membrane_row_rt(const TriGeom &g, int I, double Km[4][2][2])
{
    const double s = 1.0 / (2.0 * g.area);
    const double px[3] = {g.y23 * s, g.y31 * s, g.y12 * s};
    const double py[3] = {-g.x23 * s, -g.x31 * s, -g.x12 * s};
    const double f = c_el.thickness * g.area;
    const double pxI = pick3(I, px), pyI = pick3(I, py);
#pragma unroll
    for (int j = 0; j < 3; j++) {
        Km[j][0][0] = Km[j][0][1] = Km[j][1][0] = Km[j][1][1] = 0.0;
        membrane_accum(f, pxI, pyI, px[j], py[j], Km[j]);
    }
}

__device__ __forceinline__ void quad_membrane_row_rt(const QuadGeom &g, int I, double Km[4][2][2])
{
    const double root = 0.57735026918962584;
    const bool quirk = (c_el.quirks & FS_Q_DETLU) != 0;
#pragma unroll
    for (int j = 0; j < 4; j++) Km[j][0][0] = Km[j][0][1] = Km[j][1][0] = Km[j][1][1] = 0.0;
#pragma unroll QUAD_GP_UNROLL
    for (int gp = 0; gp < 4; gp++) {
        const double r = (gp & 2) ? -root : root;
        const double s = (gp & 1) ? -root : root;
        const double dr[4] = {-0.25 * (1 - s), 0.25 * (1 - s), 0.25 * (1 + s), -0.25 * (1 + s)};
        const double ds[4] = {-0.25 * (1 - r), -0.25 * (1 + r), 0.25 * (1 + r), 0.25 * (1 - r)};
        double j00 = 0, j01 = 0, j10 = 0, j11 = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            j00 += dr[k] * g.lx[k];
            j01 += dr[k] * g.ly[k];
            j10 += ds[k] * g.lx[k];
            j11 += ds[k] * g.ly[k];
        }
        Lu2State st = {false, false};
        const double det = det2_quirk(j00, j01, j10, j11, st, quirk);
        const double di = 1.0 / det;
        const double b00 = j11 * di, b01 = -j01 * di, b12 = -j10 * di, b13 = j00 * di;
        double px[4], py[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            px[k] = b00 * dr[k] + b01 * ds[k];
            py[k] = b12 * dr[k] + b13 * ds[k];
        }
        const double f = det * c_el.thickness;
        const MembraneRow mr = membrane_row_coeffs(f, pick4(I, px), pick4(I, py));
#pragma unroll
        for (int j = 0; j < 4; j++) membrane_accum_row(mr, px[j], py[j], Km[j]);
    }
}

// ---------------------------------------------------------------------------------------------
// Specht triangle, fs.cpp:551-603 + evalBTri fs.cpp:698-891.
// The reference spells out all 27 entries of B; the repeated brackets are S*, R*, T* here.
// ---------------------------------------------------------------------------------------------
struct TriGp {
    // entries of B that do not multiply an edge vector (columns 0, 3, 6 of each row)
    double w0[3], w1[3], w2[3];
    double S1, S2, S3, S4, R1, R2, R3, R4, T1, T2, T3, T4, T5, T6;
};

__device__ __forceinline__ void tri_gp_terms(double L1, double L2, double mu1, double mu2,
                                             double mu3, TriGp &t)
{
    const double L3 = 1 - L1 - L2;
    const double p1 = 1 + 3 * mu1, p2 = 1 + 3 * mu2, p3 = 1 + 3 * mu3;
    const double q3 = 1 - 3 * mu3, m2 = -1 + 3 * mu2, n3 = -1 - 3 * mu3;
    const double a = 3 * (1 - mu3) * L1 - p3 * L2 + p3 * L3;
    const double b = 3 * (1 - mu2) * L3 - p2 * L1 + p2 * L2;
    const double c = 3 * (1 - mu1) * L2 - p1 * L3 + p1 * L1;
    t.S1 = -2 + 6 * L1 + 4 * L2 - L2 * b - 4 * L2 * L3 + 4 * L1 * L2;
    t.S2 = 2 * L2 - L2 * a + L2 * L3 * 2 * q3 - L1 * L2 * 2 * q3;
    t.S3 = -L2 * c + L2 * L3 * 2 * p1 - L1 * L2 * 2 * p1;
    t.S4 = -4 + 6 * L1 + 4 * L2 - L2 * b - 4 * L2 * L3 + 4 * L1 * L2;
    t.w0[0] = 6 + L2 * (-4 - 2 * a) + 4 * q3 * (L2 * L3 - L1 * L2) - 12 * L1 + 2 * L2 * b +
              8 * (L2 * L3 - L1 * L2);
    t.w0[1] = -2 * L2 * c + 4 * p1 * (L2 * L3 - L1 * L2) - 4 * L2 + 2 * L2 * a +
              4 * q3 * (-L2 * L3 + L1 * L2);
    t.w0[2] = -6 + 12 * L1 + 8 * L2 - 2 * L2 * b + 8 * (L1 * L2 - L2 * L3) + 2 * L2 * c +
              4 * p1 * (L1 * L2 - L2 * L3);
    t.R1 = 2 * L1 - L1 * b + L1 * L3 * 2 * m2 - L1 * L2 * 2 * m2;
    t.R2 = -L1 * a + L1 * L3 * 2 * n3 - L1 * L2 * 2 * n3;
    t.R3 = -6 * L2 + 2 - 2 * L1 - L1 * c + 4 * L3 * L1 - 4 * L1 * L2;
    t.R4 = -6 * L2 + 4 - 2 * L1 - L1 * c + 4 * L3 * L1 - 4 * L1 * L2;
    t.w1[0] = -2 * L1 * a + 2 * L1 * L3 * 2 * n3 - 2 * L1 * L2 * 2 * n3 - 4 * L1 + 2 * L1 * b -
              2 * L1 * L3 * 2 * m2 + 2 * L1 * L2 * 2 * m2;
    t.w1[1] = 6 - 12 * L2 - 4 * L1 - 2 * L1 * c + 8 * L3 * L1 - 8 * L1 * L2 + 2 * L1 * a -
              2 * L1 * L3 * 2 * n3 + 2 * L1 * L2 * 2 * n3;
    t.w1[2] = -6 + 8 * L1 - 2 * L1 * b + 2 * L1 * L3 * 2 * m2 - 2 * L1 * L2 * 2 * m2 + 12 * L2 +
              2 * L1 * c - 8 * L3 * L1 + 8 * L1 * L2;
    t.T1 = -1 + 4 * L1 + 2 * L2 + 0.5 * L3 * b - 0.5 * L2 * b + 0.5 * L2 * L3 * 2 * m2 -
           0.5 * L1 * b - 0.5 * L1 * L2 * 2 * m2 - 2 * L3 * L1 + 2 * L1 * L2;
    t.T2 = 2 * L1 + 0.5 * L3 * a - 0.5 * L2 * a + 0.5 * L2 * L3 * 2 * n3 - 0.5 * L1 * a -
           0.5 * L1 * L2 * 2 * n3 + 0.5 * L1 * L3 * 2 * q3 - 0.5 * L1 * L2 * 2 * q3;
    t.T3 = t.T2 - 1;
    t.T4 = -2 * L2 + 0.5 * L3 * c - 0.5 * L2 * c + 2 * L2 * L3 - 0.5 * L1 * c - 2 * L1 * L2 +
           0.5 * L1 * L3 * 2 * p1 - 0.5 * L1 * L2 * 2 * p1;
    t.T5 = t.T4 + 1;
    t.T6 = t.T1 - 1;
    t.w2[0] = 2 - 4 * L1 + L3 * a - L2 * a + L2 * L3 * 2 * n3 - L1 * a - L1 * L2 * 2 * n3 +
              L1 * L3 * 2 * q3 - L1 * L2 * 2 * q3 - 4 * L2 - L3 * b + L2 * b - L2 * L3 * 2 * m2 +
              L1 * b + L1 * L2 * 2 * m2 + 4 * L3 * L1 - 4 * L1 * L2;
    t.w2[1] = 2 - 4 * L2 + L3 * c - L2 * c + 4 * L2 * L3 - L1 * c - 4 * L1 * L2 +
              L1 * L3 * 2 * p1 - L1 * L2 * 2 * p1 - 4 * L1 - L3 * a + L2 * a + L1 * a -
              L2 * L3 * 2 * n3 + L1 * L2 * 2 * n3 - L1 * L3 * 2 * q3 + L1 * L2 * 2 * q3;
    t.w2[2] = -4 + 8 * L1 + 8 * L2 + L3 * b - L2 * b + L2 * L3 * 2 * m2 - L1 * b -
              L1 * L2 * 2 * m2 - 4 * L3 * L1 + 8 * L1 * L2 - L3 * c + L2 * c - 4 * L2 * L3 +
              L1 * c - L1 * L3 * 2 * p1 + L1 * L2 * 2 * p1;
}

// columns 3J..3J+2 of B at one quadrature point; Bc[row][col]; row 2 already doubled (fs.cpp:889)
template <int J>
__device__ __forceinline__ void tri_bcols(const TriGeom &g, const TriGp &t, double Bc[3][3])
{
    // node J pairs two edge vectors with two brackets per row:
    //   J=0: (31 with S1/R1/T1, 12 with S2/R2/T2)   J=1: (12 with S2/R2/T3, 23 with S3/R3/T4)
    //   J=2: (23 with S3/R4/T5, 31 with S4/R1/T6)
    double ex1, ey1, ex2, ey2, s1, s2, r1, r2, q1, q2;
    if (J == 0) {
        ex1 = g.x31; ey1 = g.y31; ex2 = g.x12; ey2 = g.y12;
        s1 = t.S1; s2 = t.S2; r1 = t.R1; r2 = t.R2; q1 = t.T1; q2 = t.T2;
    } else if (J == 1) {
        ex1 = g.x12; ey1 = g.y12; ex2 = g.x23; ey2 = g.y23;
        s1 = t.S2; s2 = t.S3; r1 = t.R2; r2 = t.R3; q1 = t.T3; q2 = t.T4;
    } else {
        ex1 = g.x23; ey1 = g.y23; ex2 = g.x31; ey2 = g.y31;
        s1 = t.S3; s2 = t.S4; r1 = t.R4; r2 = t.R1; q1 = t.T5; q2 = t.T6;
    }
    Bc[0][0] = t.w0[J];
    Bc[0][1] = -ey1 * s1 - ey2 * s2;
    Bc[0][2] = ex1 * s1 + ex2 * s2;
    Bc[1][0] = t.w1[J];
    Bc[1][1] = -ey1 * r1 - ey2 * r2;
    Bc[1][2] = ex1 * r1 + ex2 * r2;
    Bc[2][0] = 2.0 * t.w2[J];
    Bc[2][1] = 2.0 * (-ey1 * q1 - ey2 * q2);
    Bc[2][2] = 2.0 * (ex1 * q1 + ex2 * q2);
}

template <int I>
__device__ __forceinline__ void tri_plate_row(const TriGeom &g, double Kp[3][3][3])
{
    const double C0 = g.x12 * g.x12 + g.y12 * g.y12;  // fs.cpp:566-568
    const double C1 = g.x31 * g.x31 + g.y31 * g.y31;
    const double C2 = g.x23 * g.x23 + g.y23 * g.y23;
    const double mu1 = (C0 - C1) / C2, mu2 = (C2 - C0) / C1, mu3 = (C1 - C2) / C0;  // fs.cpp:702-704
    const double sc = 1.0 / (4.0 * g.area * g.area);
    double Y[3][3];  // fs.cpp:578-588
    Y[0][0] = g.y23 * g.y23 * sc;
    Y[0][1] = g.y31 * g.y31 * sc;
    Y[0][2] = g.y23 * g.y31 * sc;
    Y[1][0] = g.x23 * g.x23 * sc;
    Y[1][1] = g.x31 * g.x31 * sc;
    Y[1][2] = g.x31 * g.x23 * sc;
    Y[2][0] = -2.0 * g.x23 * g.y23 * sc;
    Y[2][1] = ((c_el.quirks & FS_Q_Y21) ? -2.0 * g.x31 * g.x31 : -2.0 * g.x31 * g.y31) * sc;
    Y[2][2] = (-g.x23 * g.y31 - g.x31 * g.y23) * sc;
    const double d11 = c_el.dp11, d12 = c_el.dp12, d33 = c_el.dp33;
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
        for (int r = 0; r < 3; r++) Kp[j][r][0] = Kp[j][r][1] = Kp[j][r][2] = 0.0;

#pragma unroll
    for (int gp = 0; gp < 3; gp++) {
        const double L1 = (gp == 1) ? 2.0 / 3.0 : 1.0 / 6.0;  // fs.cpp:560-562
        const double L2 = (gp == 2) ? 2.0 / 3.0 : 1.0 / 6.0;
        TriGp t;
        tri_gp_terms(L1, L2, mu1, mu2, mu3, t);
        // E = Dp * (Y * B_I)  (3x3)
        double Bc[3][3], M[3][3], E[3][3];
        tri_bcols<I>(g, t, Bc);
#pragma unroll
        for (int k = 0; k < 3; k++)
#pragma unroll
            for (int c = 0; c < 3; c++) M[k][c] = Y[k][0] * Bc[0][c] + Y[k][1] * Bc[1][c] + Y[k][2] * Bc[2][c];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            E[0][c] = d11 * M[0][c] + d12 * M[1][c];
            E[1][c] = d12 * M[0][c] + d11 * M[1][c];
            E[2][c] = d33 * M[2][c];
        }
#define FS_TRI_ACC(J)                                                                              \
    {                                                                                              \
        tri_bcols<J>(g, t, Bc);                                                                    \
        _Pragma("unroll") for (int k = 0; k < 3; k++) _Pragma("unroll") for (int c = 0; c < 3; c++) \
            M[k][c] = Y[k][0] * Bc[0][c] + Y[k][1] * Bc[1][c] + Y[k][2] * Bc[2][c];               \
        _Pragma("unroll") for (int r = 0; r < 3; r++) _Pragma("unroll") for (int c = 0; c < 3; c++) \
            Kp[J][r][c] += (E[0][r] * M[0][c] + E[1][r] * M[1][c] + E[2][r] * M[2][c]) * (1.0 / 6.0); \
    }
        FS_TRI_ACC(0)
        FS_TRI_ACC(1)
        FS_TRI_ACC(2)
#undef FS_TRI_ACC
    }
    const double f = 2.0 * g.area;  // fs.cpp:602
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int c = 0; c < 3; c++) Kp[j][r][c] *= f;
}

__device__ __forceinline__ void tri_plate_row_rt(const TriGeom &g, int I, double Kp[4][3][3])
{
    const double C0 = g.x12 * g.x12 + g.y12 * g.y12;
    const double C1 = g.x31 * g.x31 + g.y31 * g.y31;
    const double C2 = g.x23 * g.x23 + g.y23 * g.y23;
    const double mu1 = (C0 - C1) / C2, mu2 = (C2 - C0) / C1, mu3 = (C1 - C2) / C0;
    const double sc = 1.0 / (4.0 * g.area * g.area);
    double Y[3][3];
    Y[0][0] = g.y23 * g.y23 * sc;
    Y[0][1] = g.y31 * g.y31 * sc;
    Y[0][2] = g.y23 * g.y31 * sc;
    Y[1][0] = g.x23 * g.x23 * sc;
    Y[1][1] = g.x31 * g.x31 * sc;
    Y[1][2] = g.x31 * g.x23 * sc;
    Y[2][0] = -2.0 * g.x23 * g.y23 * sc;
    Y[2][1] = ((c_el.quirks & FS_Q_Y21) ? -2.0 * g.x31 * g.x31 : -2.0 * g.x31 * g.y31) * sc;
    Y[2][2] = (-g.x23 * g.y31 - g.x31 * g.y23) * sc;
    const double d11 = c_el.dp11, d12 = c_el.dp12, d33 = c_el.dp33;
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
        for (int r = 0; r < 3; r++) Kp[j][r][0] = Kp[j][r][1] = Kp[j][r][2] = 0.0;
#pragma unroll
    for (int gp = 0; gp < 3; gp++) {
        const double L1 = (gp == 1) ? 2.0 / 3.0 : 1.0 / 6.0;
        const double L2 = (gp == 2) ? 2.0 / 3.0 : 1.0 / 6.0;
        TriGp t;
        tri_gp_terms(L1, L2, mu1, mu2, mu3, t);
        double M[3][3][3];  // M[j] = Y * B_j
#define FS_TRI_M(J)                                                                                \
    {                                                                                              \
        double Bc[3][3];                                                                           \
        tri_bcols<J>(g, t, Bc);                                                                    \
        _Pragma("unroll") for (int k = 0; k < 3; k++) _Pragma("unroll") for (int c = 0; c < 3; c++) \
            M[J][k][c] = Y[k][0] * Bc[0][c] + Y[k][1] * Bc[1][c] + Y[k][2] * Bc[2][c];            \
    }
        FS_TRI_M(0)
        FS_TRI_M(1)
        FS_TRI_M(2)
#undef FS_TRI_M
        double E[3][3];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const double m0 = sel_f64(I == 0, M[0][0][c], sel_f64(I == 1, M[1][0][c], M[2][0][c]));
            const double m1 = sel_f64(I == 0, M[0][1][c], sel_f64(I == 1, M[1][1][c], M[2][1][c]));
            const double m2 = sel_f64(I == 0, M[0][2][c], sel_f64(I == 1, M[1][2][c], M[2][2][c]));
            E[0][c] = (d11 * m0 + d12 * m1) * (1.0 / 6.0);   // the Gauss weight rides on E (fs.cpp:560-562)
            E[1][c] = (d12 * m0 + d11 * m1) * (1.0 / 6.0);
            E[2][c] = d33 * m2 * (1.0 / 6.0);
        }
#pragma unroll
        for (int j = 0; j < 3; j++)
#pragma unroll
            for (int r = 0; r < 3; r++)
#pragma unroll
                for (int c = 0; c < 3; c++)
                    Kp[j][r][c] = fma(E[2][r], M[j][2][c], fma(E[1][r], M[j][1][c], fma(E[0][r], M[j][0][c], Kp[j][r][c])));
    }
    const double f = 2.0 * g.area;
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int c = 0; c < 3; c++) Kp[j][r][c] *= f;
}

// ---------------------------------------------------------------------------------------------
// DKQ, fs.cpp:604-687 + evalBQuad fs.cpp:901-990
// ---------------------------------------------------------------------------------------------
struct QuadH {
    double a[4], b[4], c[4], d[4], e[4];  // Hcoeffs rows, fs.cpp:613-621
};

// columns 3K..3K+2 of B at (xi, eta): node K couples mid-side K (s) and mid-side K-1 (p)
template <int K>
__device__ __forceinline__ void quad_bcols(const QuadH &h, double xi, double eta, double i00,
                                           double i01, double i10, double i11, double Bc[3][3])
{
    // serendipity derivatives, fs.cpp:907-923 (corner K and the two adjacent mid-side nodes)
    double Nxk, Nek;
    if (K == 0) { Nxk = 0.25 * (2.0 * xi + eta) * (1.0 - eta); Nek = 0.25 * (2.0 * eta + xi) * (1.0 - xi); }
    else if (K == 1) { Nxk = 0.25 * (2.0 * xi - eta) * (1.0 - eta); Nek = 0.25 * (2.0 * eta - xi) * (1.0 + xi); }
    else if (K == 2) { Nxk = 0.25 * (2.0 * xi + eta) * (1.0 + eta); Nek = 0.25 * (2.0 * eta + xi) * (1.0 + xi); }
    else { Nxk = 0.25 * (2.0 * xi - eta) * (1.0 + eta); Nek = 0.25 * (2.0 * eta - xi) * (1.0 - xi); }
    const double Nx_mid[4] = {-xi * (1.0 - eta), 0.5 * (1.0 - eta * eta), -xi * (1.0 + eta), -0.5 * (1.0 - eta * eta)};
    const double Ne_mid[4] = {-0.5 * (1.0 - xi * xi), -eta * (1.0 + xi), 0.5 * (1.0 - xi * xi), -eta * (1.0 - xi)};
    constexpr int S = K, P = (K + 3) & 3;
    const double nxs = Nx_mid[S], nxp = Nx_mid[P], nes = Ne_mid[S], nep = Ne_mid[P];
    // Hx_xi, Hx_eta, Hy_xi, Hy_eta for the three DOFs of node K, fs.cpp:931-981
    double hxx[3], hxe[3], hyx[3], hye[3];
    hxx[0] = 1.5 * (h.a[S] * nxs - h.a[P] * nxp);
    hxx[1] = h.b[S] * nxs + h.b[P] * nxp;
    hxx[2] = Nxk - h.c[S] * nxs - h.c[P] * nxp;
    hyx[0] = 1.5 * (h.d[S] * nxs - h.d[P] * nxp);
    hyx[1] = -Nxk + h.e[S] * nxs + h.e[P] * nxp;
    hyx[2] = -hxx[1];
    hxe[0] = 1.5 * (h.a[S] * nes - h.a[P] * nep);
    hxe[1] = h.b[S] * nes + h.b[P] * nep;
    hxe[2] = Nek - h.c[S] * nes - h.c[P] * nep;
    hye[0] = 1.5 * (h.d[S] * nes - h.d[P] * nep);
    hye[1] = -Nek + h.e[S] * nes + h.e[P] * nep;
    hye[2] = -hxe[1];
#pragma unroll
    for (int c = 0; c < 3; c++) {  // fs.cpp:984-989
        Bc[0][c] = i00 * hxx[c] + i01 * hxe[c];
        Bc[1][c] = i10 * hyx[c] + i11 * hye[c];
        Bc[2][c] = i00 * hyx[c] + i01 * hye[c] + i10 * hxx[c] + i11 * hxe[c];
    }
}

template <int I>
__device__ __forceinline__ void quad_plate_row(const QuadGeom &g, double Kp[4][3][3])
{
    QuadH h;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const double dx = g.dx[k], dy = g.dy[k];
        const double sl = dx * dx + dy * dy;  // fs.cpp:608-611
        h.a[k] = -dx / sl;
        h.b[k] = 0.75 * dx * dy / sl;
        h.c[k] = (0.25 * dx * dx - 0.5 * dy * dy) / sl;
        h.d[k] = -dy / sl;
        h.e[k] = (0.25 * dy * dy - 0.5 * dx * dx) / sl;
    }
    const double d11 = c_el.dp11, d12 = c_el.dp12, d33 = c_el.dp33;
    const double root = 0.57735026918962584;
    const bool quirk = (c_el.quirks & FS_Q_DETLU) != 0;
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
        for (int r = 0; r < 3; r++) Kp[j][r][0] = Kp[j][r][1] = Kp[j][r][2] = 0.0;
    Lu2State st = {false, false};  // J lives across the GP loop: fs.cpp:633
#pragma unroll
    for (int gp = 0; gp < 4; gp++) {
        const double r = (gp & 2) ? -root : root;
        const double s = (gp & 1) ? -root : root;
        double j00 = 0.25 * ((g.dx[0] + g.dx[2]) * s - g.dx[0] + g.dx[2]);  // fs.cpp:641-645
        double j01 = 0.25 * ((g.dy[0] + g.dy[2]) * s - g.dy[0] + g.dy[2]);
        double j10 = 0.25 * ((g.dx[0] + g.dx[2]) * r - g.dx[1] + g.dx[3]);
        double j11 = 0.25 * ((g.dy[0] + g.dy[2]) * r - g.dy[1] + g.dy[3]);
        const double det = det2_quirk(j00, j01, j10, j11, st, quirk);  // fs.cpp:652
        const double di = 1.0 / det;
        const double i00 = j11 * di, i01 = -j01 * di, i10 = -j10 * di, i11 = j00 * di;  // fs.cpp:656-660
        double Bc[3][3], E[3][3];
        quad_bcols<I>(h, r, s, i00, i01, i10, i11, Bc);
#pragma unroll
        for (int c = 0; c < 3; c++) {  // E = Dp * B_I
            E[0][c] = d11 * Bc[0][c] + d12 * Bc[1][c];
            E[1][c] = d12 * Bc[0][c] + d11 * Bc[1][c];
            E[2][c] = d33 * Bc[2][c];
        }
#define FS_QUAD_ACC(J)                                                                             \
    {                                                                                              \
        quad_bcols<J>(h, r, s, i00, i01, i10, i11, Bc);                                            \
        _Pragma("unroll") for (int rr = 0; rr < 3; rr++) _Pragma("unroll") for (int c = 0; c < 3; c++) \
            Kp[J][rr][c] += (E[0][rr] * Bc[0][c] + E[1][rr] * Bc[1][c] + E[2][rr] * Bc[2][c]) * det; \
    }
        FS_QUAD_ACC(0)
        FS_QUAD_ACC(1)
        FS_QUAD_ACC(2)
        FS_QUAD_ACC(3)
#undef FS_QUAD_ACC
    }
}

// Shape-function derivatives of corner K and of its two adjacent mid-side nodes at the four Gauss points,
// fs.cpp:907-923.  With a run-time K these would be selections among compile-time constants, which the
// compiler turns into divergent branch trees; the host tabulates them once (quad_gp_table) and the lanes of
// the gather kernel read their six numbers per Gauss point from a shared-memory copy.
struct QuadGpTab {
    double v[4][4][6];  // [gp][K] -> Nx_K, Ne_K, Nx_mid[K], Nx_mid[K-1], Ne_mid[K], Ne_mid[K-1]
};

inline void quad_gp_table(QuadGpTab &t)
{
    const double root = 0.57735026918962584;  // sqrt(1.0/3.0), fs.cpp:472
    for (int gp = 0; gp < 4; gp++) {
        const double xi = (gp & 2) ? -root : root, eta = (gp & 1) ? -root : root;
        const double Nx_c[4] = {0.25 * (2.0 * xi + eta) * (1.0 - eta), 0.25 * (2.0 * xi - eta) * (1.0 - eta),
                                0.25 * (2.0 * xi + eta) * (1.0 + eta), 0.25 * (2.0 * xi - eta) * (1.0 + eta)};
        const double Ne_c[4] = {0.25 * (2.0 * eta + xi) * (1.0 - xi), 0.25 * (2.0 * eta - xi) * (1.0 + xi),
                                0.25 * (2.0 * eta + xi) * (1.0 + xi), 0.25 * (2.0 * eta - xi) * (1.0 - xi)};
        const double Nx_mid[4] = {-xi * (1.0 - eta), 0.5 * (1.0 - eta * eta), -xi * (1.0 + eta), -0.5 * (1.0 - eta * eta)};
        const double Ne_mid[4] = {-0.5 * (1.0 - xi * xi), -eta * (1.0 + xi), 0.5 * (1.0 - xi * xi), -eta * (1.0 - xi)};
        for (int K = 0; K < 4; K++) {
            const int P = (K + 3) & 3;
            double *o = t.v[gp][K];
            o[0] = Nx_c[K]; o[1] = Ne_c[K]; o[2] = Nx_mid[K]; o[3] = Nx_mid[P]; o[4] = Ne_mid[K]; o[5] = Ne_mid[P];
        }
    }
}

// columns of node K chosen at run time: tb = the six tabulated numbers of (gp, K); the side coefficients are
// register selects
struct QuadHK {
    double aS, aP, bS, bP, cS, cP, dS, dP, eS, eP;  // Hcoeffs of side K (after node K) and side K-1 (before it)
};

__device__ __forceinline__ void quad_pick_sides(const QuadH &h, int K, QuadHK &k)
{
    const int P = (K + 3) & 3;
    k.aS = pick4(K, h.a); k.aP = pick4(P, h.a); k.bS = pick4(K, h.b); k.bP = pick4(P, h.b);
    k.cS = pick4(K, h.c); k.cP = pick4(P, h.c); k.dS = pick4(K, h.d); k.dP = pick4(P, h.d);
    k.eS = pick4(K, h.e); k.eP = pick4(P, h.e);
}

__device__ __forceinline__ void quad_bcols_rt(const QuadHK &k, const double *tb, double i00, double i01,
                                              double i10, double i11, double Bc[3][3])
{
    const double aS = k.aS, aP = k.aP, bS = k.bS, bP = k.bP, cS = k.cS, cP = k.cP, dS = k.dS, dP = k.dP, eS = k.eS, eP = k.eP;
    const double2 t01 = *reinterpret_cast<const double2 *>(tb), t23 = *reinterpret_cast<const double2 *>(tb + 2),
                  t45 = *reinterpret_cast<const double2 *>(tb + 4);
    const double Nxk = t01.x, Nek = t01.y, nxs = t23.x, nxp = t23.y, nes = t45.x, nep = t45.y;
    double hxx[3], hxe[3], hyx[3], hye[3];
    hxx[0] = 1.5 * (aS * nxs - aP * nxp);
    hxx[1] = bS * nxs + bP * nxp;
    hxx[2] = Nxk - cS * nxs - cP * nxp;
    hyx[0] = 1.5 * (dS * nxs - dP * nxp);
    hyx[1] = -Nxk + eS * nxs + eP * nxp;
    hyx[2] = -hxx[1];
    hxe[0] = 1.5 * (aS * nes - aP * nep);
    hxe[1] = bS * nes + bP * nep;
    hxe[2] = Nek - cS * nes - cP * nep;
    hye[0] = 1.5 * (dS * nes - dP * nep);
    hye[1] = -Nek + eS * nes + eP * nep;
    hye[2] = -hxe[1];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        Bc[0][c] = i00 * hxx[c] + i01 * hxe[c];
        Bc[1][c] = i10 * hyx[c] + i11 * hye[c];
        Bc[2][c] = i00 * hyx[c] + i01 * hye[c] + i10 * hxx[c] + i11 * hxe[c];
    }
}

// qtab: shared-memory copy of the QuadGpTab
__device__ __forceinline__ void quad_plate_row_rt(const QuadGeom &g, int I, const double *qtab, double Kp[4][3][3])
{
    QuadH h;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const double dx = g.dx[k], dy = g.dy[k];
        const double si = 1.0 / (dx * dx + dy * dy);  // one reciprocal per side instead of five divisions
        h.a[k] = -dx * si;
        h.b[k] = 0.75 * dx * dy * si;
        h.c[k] = (0.25 * dx * dx - 0.5 * dy * dy) * si;
        h.d[k] = -dy * si;
        h.e[k] = (0.25 * dy * dy - 0.5 * dx * dx) * si;
    }
    QuadHK hk;
    quad_pick_sides(h, I, hk);
    const double d11 = c_el.dp11, d12 = c_el.dp12, d33 = c_el.dp33;
    const double root = 0.57735026918962584;
    const bool quirk = (c_el.quirks & FS_Q_DETLU) != 0;
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
        for (int r = 0; r < 3; r++) Kp[j][r][0] = Kp[j][r][1] = Kp[j][r][2] = 0.0;
    Lu2State st = {false, false};
#pragma unroll QUAD_GP_UNROLL
    for (int gp = 0; gp < 4; gp++) {
        const double r = (gp & 2) ? -root : root;
        const double s = (gp & 1) ? -root : root;
        double j00 = 0.25 * ((g.dx[0] + g.dx[2]) * s - g.dx[0] + g.dx[2]);
        double j01 = 0.25 * ((g.dy[0] + g.dy[2]) * s - g.dy[0] + g.dy[2]);
        double j10 = 0.25 * ((g.dx[0] + g.dx[2]) * r - g.dx[1] + g.dx[3]);
        double j11 = 0.25 * ((g.dy[0] + g.dy[2]) * r - g.dy[1] + g.dy[3]);
        const double det = det2_quirk(j00, j01, j10, j11, st, quirk);
        const double di = 1.0 / det;
        const double i00 = j11 * di, i01 = -j01 * di, i10 = -j10 * di, i11 = j00 * di;
        double Bc[3][3], E[3][3];
        quad_bcols_rt(hk, qtab + (gp * 4 + I) * 6, i00, i01, i10, i11, Bc);
        const double e11 = d11 * det, e12 = d12 * det, e33 = d33 * det;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            E[0][c] = fma(e12, Bc[1][c], e11 * Bc[0][c]);
            E[1][c] = fma(e11, Bc[1][c], e12 * Bc[0][c]);
            E[2][c] = e33 * Bc[2][c];
        }
#define FS_QUAD_ACC_RT(J)                                                                          \
    {                                                                                              \
        quad_bcols<J>(h, r, s, i00, i01, i10, i11, Bc);                                            \
        _Pragma("unroll") for (int rr = 0; rr < 3; rr++) _Pragma("unroll") for (int c = 0; c < 3; c++) \
            Kp[J][rr][c] = fma(E[2][rr], Bc[2][c], fma(E[1][rr], Bc[1][c], fma(E[0][rr], Bc[0][c], Kp[J][rr][c]))); \
    }
        FS_QUAD_ACC_RT(0)
        FS_QUAD_ACC_RT(1)
        FS_QUAD_ACC_RT(2)
        FS_QUAD_ACC_RT(3)
#undef FS_QUAD_ACC_RT
    }
}

// ---------------------------------------------------------------------------------------------
// superposition (fs.cpp:1011-1051) + rotation to global axes (fs.cpp:1084-1102) of ONE node block.
// Local block: rows/cols (u v | w tx ty | tz) = membrane 2x2, plate 3x3, drilling scalar.
// G = Tt^T K Tt with Tt = diag(T, T); computed as Tt^T (K Tt) like the reference.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void rotate_block(const double T[3][3], const double m[2][2],
                                             const double p[3][3], double G[6][6])
{
    // drilling stiffness: max of the five diagonal entries / 1000 -- on every block (i,j)
    double d = fmax(m[0][0], m[1][1]);
    d = fmax(d, p[0][0]);
    d = fmax(d, p[1][1]);
    d = fmax(d, p[2][2]);
    d /= 1000.0;
    double KT[3][3];
    // translations x translations: K = [[m00 m01 0][m10 m11 0][0 0 pww]]
#pragma unroll
    for (int c = 0; c < 3; c++) {
        KT[0][c] = m[0][0] * T[0][c] + m[0][1] * T[1][c];
        KT[1][c] = m[1][0] * T[0][c] + m[1][1] * T[1][c];
        KT[2][c] = p[0][0] * T[2][c];
    }
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) G[r][c] = T[0][r] * KT[0][c] + T[1][r] * KT[1][c] + T[2][r] * KT[2][c];
    // translations x rotations: K = [[0 0 0][0 0 0][pwx pwy 0]]
#pragma unroll
    for (int c = 0; c < 3; c++) KT[2][c] = p[0][1] * T[0][c] + p[0][2] * T[1][c];
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) G[r][3 + c] = T[2][r] * KT[2][c];
    // rotations x translations: K = [[0 0 pxw][0 0 pyw][0 0 0]]
#pragma unroll
    for (int c = 0; c < 3; c++) {
        KT[0][c] = p[1][0] * T[2][c];
        KT[1][c] = p[2][0] * T[2][c];
    }
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) G[3 + r][c] = T[0][r] * KT[0][c] + T[1][r] * KT[1][c];
    // rotations x rotations: K = [[pxx pxy 0][pyx pyy 0][0 0 d]]
#pragma unroll
    for (int c = 0; c < 3; c++) {
        KT[0][c] = p[1][1] * T[0][c] + p[1][2] * T[1][c];
        KT[1][c] = p[2][1] * T[0][c] + p[2][2] * T[1][c];
        KT[2][c] = d * T[2][c];
    }
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) G[3 + r][3 + c] = T[0][r] * KT[0][c] + T[1][r] * KT[1][c] + T[2][r] * KT[2][c];
}

// node row I of a triangle: the three rotated 6x6 blocks go to sink.block(j, G)
template <int I, class Sink>
__device__ __forceinline__ void tri_row_blocks(const double X[9], Sink &sink)
{
    TriGeom g;
    tri_geom(X, g);
    double Km[3][2][2], Kp[3][3][3];
    tri_membrane_row<I>(g, Km);
    tri_plate_row<I>(g, Kp);
#pragma unroll
    for (int j = 0; j < 3; j++) {
        double G[6][6];
        rotate_block(g.T, Km[j], Kp[j], G);
        sink.block(j, G);
    }
}

template <int I, class Sink>
__device__ __forceinline__ void quad_row_blocks(const double X[12], Sink &sink)
{
    QuadGeom g;
    quad_geom(X, g);
    double Km[4][2][2], Kp[4][3][3];
    quad_membrane_row<I>(g, Km);
    quad_plate_row<I>(g, Kp);
#pragma unroll
    for (int j = 0; j < 4; j++) {
        double G[6][6];
        rotate_block(g.T, Km[j], Kp[j], G);
        sink.block(j, G);
    }
}

}  // namespace fs

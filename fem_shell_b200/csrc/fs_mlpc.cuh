// fs_mlpc.cuh -- data of the multilevel rigid-body-mode preconditioner (FS_PC_MLRBM): a smoothed-aggregation
// V/W-cycle whose coarse levels live on nested regular lattices laid over the bounding box of the mesh.
//
// Not a piece of the reference: fem-shell hands the Krylov solve to PETSc (fs.cpp:138) and its documented
// options stop at point/block Jacobi (doc/implementation.tex:68-72).  With those, CG on a clamped plate
// needs O(n^2) iterations for n nodes per side (464 846 for BASELINE configs[1]), which puts
// time-to-solution of the 96 M-DOF target out of reach on any hardware.  This preconditioner keeps the
// Krylov method (CG on the same matrix, same tolerance, same converged displacements) and replaces
// z = D^-1 r by one multigrid cycle:
//
//   level 0      the mesh: matrix = the assembled block CSR, smoother = damped 6x6 block Jacobi
//   level l >= 1 a lattice of cells of size H_l = 3^(l-1) H_1 (H_1 = three element widths per active
//                dimension; dimensions of zero extent carry one layer).  A cell is one aggregate; its six
//                unknowns are the rigid-body modes about the cell centre, u = t + w x (x - c), theta = w,
//                which is exact for the 6-DOF flat-shell nodes (fs.cpp:1061-1110 rotates every node block
//                to global axes).  Tentative prolongator P_t = these modes (rows of Dirichlet DOFs zeroed),
//                prolongator P = (I - omega D^-1 A) P_t, coarse matrix = P^T A P, a 3^d-point stencil of
//                6x6 blocks obtained by probing: cells of one colour (index mod 3 per dimension) do not
//                interact, so 3^d colours x 6 modes applications of the cycle's own transfer kernels give
//                every stencil entry.  The smoothing is applied matrix-free (one SpMV each way).
//   coarsest     at most ml_dense_points cells: dense pseudo-inverse (Gauss-Jordan on the device)
//
// Levels >= 1 are replicated on every rank; only the restricted residual of level 1 is all-reduced.
// All sums are taken in a fixed order (no atomics).  Design evidence (iteration counts of this exact
// scheme against Jacobi and against an additive hat-function variant): tools/ml_lab.py.
#pragma once
#include <cstdint>

namespace fs {

constexpr int ML_MAX_LEVELS = 14;       // lattice levels
// cells of the dense coarsest level (6 unknowns each).  A lattice visit is a chain of ~12 kernels of a few microseconds and
// a W-cycle visits level l 2^l times: on the 96 M-DOF plate the levels below 10^4 cells took 2.2 of 7.4 ms per iteration
// (profiles/r02l_ml_stage_profile_n8.json).  Ending the hierarchy at <= 400 cells (a 2400 x 2400 inverse, 46 MB, applied
// from L2 in ~8 us) instead of <= 200 removes the two deepest levels of that plate for ~20 ms more set-up.
constexpr int ML_DENSE_MAX_POINTS = 512;
constexpr int ML_DENSE_DEFAULT_POINTS = 400;

struct LatGeom {
    int np[3];      // cells per dimension (1 for an inactive dimension)
    int active[3];
    int n;          // np[0] * np[1] * np[2]
    int ns;         // stencil size 3^(active dimensions)
    double lo[3];   // corner of cell (0,0,0); the coordinate itself for an inactive dimension
    double H[3];    // cell size (0 for an inactive dimension)
};

__host__ __device__ inline void lat_unindex(const LatGeom &g, int a, int k[3])
{
    k[0] = a % g.np[0];
    const int rest = a / g.np[0];
    k[1] = rest % g.np[1];
    k[2] = rest / g.np[1];
}

__host__ __device__ inline int lat_index(const LatGeom &g, const int k[3]) { return (k[2] * g.np[1] + k[1]) * g.np[0] + k[0]; }

__host__ __device__ inline void lat_centre(const LatGeom &g, const int k[3], double c[3])
{
    for (int d = 0; d < 3; d++) c[d] = g.lo[d] + ((double)k[d] + 0.5) * g.H[d];
}

// cell containing x (clamped to the lattice)
__host__ __device__ inline int lat_cell_of(const LatGeom &g, const double x[3])
{
    int k[3];
    for (int d = 0; d < 3; d++) {
        if (g.active[d]) {
            int kk = (int)((x[d] - g.lo[d]) / g.H[d]);
            if (kk < 0) kk = 0;
            if (kk > g.np[d] - 1) kk = g.np[d] - 1;
            k[d] = kk;
        } else k[d] = 0;
    }
    return lat_index(g, k);
}

// stencil slot s -> offsets in {-1,0,1} per dimension (0 for inactive dimensions); digit i of s (base 3)
// belongs to the i-th ACTIVE dimension
__host__ __device__ inline void lat_stencil_off(const LatGeom &g, int s, int o[3])
{
    for (int d = 0; d < 3; d++) {
        if (g.active[d]) {
            o[d] = s % 3 - 1;
            s /= 3;
        } else o[d] = 0;
    }
}

__host__ __device__ inline int lat_stencil_slot(const LatGeom &g, const int o[3])
{
    int s = 0, mul = 1;
    for (int d = 0; d < 3; d++)
        if (g.active[d]) {
            s += (o[d] + 1) * mul;
            mul *= 3;
        }
    return s;
}

}  // namespace fs

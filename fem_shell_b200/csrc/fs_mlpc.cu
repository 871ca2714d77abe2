// fs_mlpc.cu -- multilevel rigid-body-mode preconditioner (FS_PC_MLRBM): lattice hierarchy, the kernels of
// one cycle, and the set-up of the coarse stencils by probing with the cycle's own transfer kernels.
// What it is and why it exists: fs_mlpc.cuh.
//
// One application z = M^-1 r (all on the context stream, fixed summation order, no atomics):
//   mesh level      k_f_smooth0       x = w D^-1 r
//                   SpMV, k_f_resid   r1 = r - A x ; t = D^-1 r1
//                   SpMV, k_f_restrict  b_1 = P_t^T (r1 - w A t)          (8 lanes per aggregate)
//                   [ncclAllReduce of b_1 when there are several ranks]
//   lattice level   k_lat_smooth0, then gamma times { k_lat_stencil<RESID>, <RSMOOTH>, k_lat_restrict,
//                   recurse, k_lat_prolong_t, k_lat_stencil<PADD> }, k_lat_stencil<POST>
//   coarsest        k_dense_matvec with the pseudo-inverse
//   mesh level      k_f_prolong_t, SpMV, k_f_prolong_add   x += t - w D^-1 A t,  t = P_t e_1
//                   SpMV, k_f_post_finish   z = x + w D^-1 (r - A x); partial r.z; advances the CG recurrence
//
// Distributed lattice levels (several ranks).  The cells of a lattice are dealt out in SLABS -- all cells sharing the
// index of the slowest-varying active dimension -- in the order of the ranks' node blocks: rank j owns the slabs from
// the one holding its first owned node up to the one before rank j+1's (bounds S_j; parent level: ceil(S_j / 3)).
// Every array keeps its full size and global cell indices; a rank computes its own slabs only and keeps `halo` slabs
// of its neighbours valid on either side (exchanged before each stencil application).  Restrictions form partial
// sums where they are computed; the one boundary slab whose owner is the neighbour travels there and is added (two
// addends: order-independent).  Below ML_DIST_MIN_CELLS cells, or with fewer than two slabs per rank, a level (and
// everything coarser) is replicated as before: the last distributed level all-reduces its restricted residual.
// The scheme needs the node blocks to follow the slab direction (true for meshGen's numbering and the first-encounter
// order on it); otherwise every level stays replicated and only the restricted level-1 residual is all-reduced.
#include <algorithm>
#include <cmath>

#include "fs_cg_device.cuh"
#include "fs_context.hpp"
#include "fs_mlpc.cuh"
#include "fs_nccl.hpp"
#include "fs_sell.cuh"

namespace fs {

static inline unsigned int nblk(int64_t n, int bs) { return (unsigned int)std::max<int64_t>(1, (n + bs - 1) / bs); }

#define FS_NCCL_ML(ctx, call)                                                                  \
    do {                                                                                       \
        ncclResult_t r__ = (call);                                                             \
        if (r__ != ncclSuccess)                                                                \
            return fs::fail(ctx, FS_ERR_COMM, std::string(#call) + ": " + fs::nccl().GetErrorString(r__)); \
    } while (0)

// nodal values (u, theta) of the rigid-body mode e = (t, w) about a centre at distance rho: u = t + w x rho
__device__ __forceinline__ void rbm_apply(const double rho[3], const double e[6], double v[6])
{
    v[0] = e[0] + e[4] * rho[2] - e[5] * rho[1];
    v[1] = e[1] + e[5] * rho[0] - e[3] * rho[2];
    v[2] = e[2] + e[3] * rho[1] - e[4] * rho[0];
    v[3] = e[3];
    v[4] = e[4];
    v[5] = e[5];
}

// transpose: y += B^T s
__device__ __forceinline__ void rbm_apply_t(const double rho[3], const double s[6], double y[6])
{
    y[0] += s[0];
    y[1] += s[1];
    y[2] += s[2];
    y[3] += s[3] + rho[1] * s[2] - rho[2] * s[1];
    y[4] += s[4] + rho[2] * s[0] - rho[0] * s[2];
    y[5] += s[5] + rho[0] * s[1] - rho[1] * s[0];
}

#define ML_RETURN_IF_DONE(state, chk) \
    if ((chk) && (state)->done) return

// ---------------------------------------------------------------------------------------------
// mesh level (thread = one DOF of an owned node unless noted; dinv = 6x6 block inverses, row-major)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double dinv_row(const double *__restrict__ dinv, int64_t t, const double *__restrict__ v6)
{
    // row (t % 6) of block (t / 6) times the node's six values
    const double2 *d = reinterpret_cast<const double2 *>(dinv + 6 * t);
    const double2 *v = reinterpret_cast<const double2 *>(v6);
    double s = 0.0;
#pragma unroll
    for (int h = 0; h < 3; h++) {
        const double2 a = d[h], b = v[h];
        s += a.x * b.x + a.y * b.y;
    }
    return s;
}

// the diagonal-block inverses of the mesh level, either as extracted (36 per node, mask = 0) or compacted to the set bits of
// `mask` (popcount(mask) per node, row-major).  The compacted sum spells out dinv_row's pairs with literal zeros for the
// skipped entries: same products, same order, same rounding.
struct DinvRef {
    const double *p;
    unsigned long long mask;
};
__device__ __forceinline__ double dinv_row(const DinvRef d, int64_t t, const double *__restrict__ v6)
{
    if (d.mask == 0) return dinv_row(d.p, t, v6);
    const int64_t node = t / 6;
    const int a = (int)(t - 6 * node);
    const unsigned row = (unsigned)(d.mask >> (6 * a)) & 63u;
    const double *e = d.p + (size_t)__popcll(d.mask) * node + __popcll(d.mask & ((1ull << (6 * a)) - 1ull));
    double s = 0.0;
    int k = 0;
#pragma unroll
    for (int h = 0; h < 3; h++) {
        const bool bx = (row >> (2 * h)) & 1u, by = (row >> (2 * h + 1)) & 1u;
        if (bx || by) {
            const double dx = bx ? e[k] : 0.0;
            k += bx ? 1 : 0;
            const double dy = by ? e[k] : 0.0;
            k += by ? 1 : 0;
            s += dx * v6[2 * h] + dy * v6[2 * h + 1];
        }
    }
    return s;
}

// d_minv (36 per node) -> the entries of `mask`, row-major; *bad is set when an entry outside the mask is not an exact zero
__global__ void k_f_dinv_compact(int64_t n_own, unsigned long long mask, const double *__restrict__ dinv, double *__restrict__ out, int *bad)
{
    const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= n_own) return;
    const int nz = __popcll(mask);
    int k = 0;
    bool off = false;
    for (int i = 0; i < 36; i++) {
        const double v = dinv[36 * (size_t)p + i];
        if ((mask >> i) & 1ull) out[(size_t)nz * p + k++] = v;
        else off |= v != 0.0;
    }
    if (off) *bad = 1;
}

__global__ void __launch_bounds__(256)
k_f_smooth0(int64_t n6, const double *__restrict__ b, const DinvRef dinv, double omega, double *__restrict__ x,
            const CgState *state, int chk)
{
    ML_RETURN_IF_DONE(state, chk);
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n6) return;
    x[t] = omega * dinv_row(dinv, t, b + 6 * (t / 6));
}

// r1 = b - q (b may be null: zero) ; t = D^-1 r1
__global__ void __launch_bounds__(192)
k_f_resid(int64_t n6, const double *__restrict__ b, const double *__restrict__ q, const DinvRef dinv,
          double *__restrict__ r1, double *__restrict__ tv, const CgState *state, int chk)
{
    ML_RETURN_IF_DONE(state, chk);
    __shared__ __align__(16) double sh[192];
    const int64_t t = blockIdx.x * (int64_t)192 + threadIdx.x;
    const bool valid = t < n6;
    double rr = 0.0;
    if (valid) {
        rr = (b ? b[t] : 0.0) - q[t];
        r1[t] = rr;
    }
    sh[threadIdx.x] = rr;
    __syncthreads();
    if (valid) tv[t] = dinv_row(dinv, t, sh + (threadIdx.x / 6) * 6);
}

// b_1[a] = sum over the owned nodes i of aggregate a of B_i^T (r1_i - w q_i); eight lanes per aggregate
__global__ void __launch_bounds__(256)
k_f_restrict(const __grid_constant__ LatGeom g, int a0, int a1, const int32_t *__restrict__ sup_ptr, const int32_t *__restrict__ sup_node,
             const double *__restrict__ xyz_own, const uint8_t *__restrict__ mask_own, const double *__restrict__ r1,
             const double *__restrict__ q, double omega, double *__restrict__ y, const CgState *state, int chk)
{
    ML_RETURN_IF_DONE(state, chk);
    const int a = a0 + (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 3), sub = threadIdx.x & 7;
    const bool valid = a < a1;  // whole groups of eight share `a`; shuffles below stay inside the group
    double acc[6] = {0, 0, 0, 0, 0, 0};
    if (valid) {
        int k[3];
        double ca[3];
        lat_unindex(g, a, k);
        lat_centre(g, k, ca);
        const int e1 = sup_ptr[a + 1];
        for (int e = sup_ptr[a] + sub; e < e1; e += 8) {
            const int p = sup_node[e];
            const double rho[3] = {xyz_own[3 * (size_t)p] - ca[0], xyz_own[3 * (size_t)p + 1] - ca[1], xyz_own[3 * (size_t)p + 2] - ca[2]};
            double rv[6], qv[6];
            load6(r1 + 6 * (size_t)p, rv);
            load6(q + 6 * (size_t)p, qv);
            const unsigned m = mask_own[p];
#pragma unroll
            for (int c = 0; c < 6; c++) rv[c] = ((m >> c) & 1u) ? 0.0 : rv[c] - omega * qv[c];
            rbm_apply_t(rho, rv, acc);
        }
    }
#pragma unroll
    for (int c = 0; c < 6; c++) {
        acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], 4);
        acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], 2);
        acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], 1);
    }
    if (valid && sub == 0) store6(y + 6 * (size_t)a, acc);
}

// t_i = B_i e[aggregate of i] for every LOCAL node (owned and halo: e is replicated, so no exchange is needed)
__global__ void __launch_bounds__(256)
k_f_prolong_t(const __grid_constant__ LatGeom g, int64_t n_local, const int32_t *__restrict__ agg, const double *__restrict__ xyz,
              const uint8_t *__restrict__ mask, const double *__restrict__ e, double *__restrict__ tv, const CgState *state, int chk)
{
    ML_RETURN_IF_DONE(state, chk);
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_local) return;
    const int a = agg[i];
    int k[3];
    double ca[3], ev[6], v[6];
    lat_unindex(g, a, k);
    lat_centre(g, k, ca);
    const double rho[3] = {xyz[3 * i] - ca[0], xyz[3 * i + 1] - ca[1], xyz[3 * i + 2] - ca[2]};
    load6(e + 6 * (size_t)a, ev);
    rbm_apply(rho, ev, v);
    const unsigned m = mask[i];
#pragma unroll
    for (int c = 0; c < 6; c++)
        if ((m >> c) & 1u) v[c] = 0.0;
    store6(tv + 6 * i, v);
}

// x (+)= t - w D^-1 q
__global__ void __launch_bounds__(256)
k_f_prolong_add(int64_t n6, const double *__restrict__ tv, const double *__restrict__ q, const DinvRef dinv,
                double omega, double *__restrict__ x, int accumulate, const CgState *state, int chk)
{
    ML_RETURN_IF_DONE(state, chk);
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n6) return;
    const double v = tv[t] - omega * dinv_row(dinv, t, q + 6 * (t / 6));
    x[t] = accumulate ? x[t] + v : v;
}

// z = x + w D^-1 (b - q) ; partial b.z ; (INIT: p = z).  The block finishing the reduction advances the CG
// recurrence: red[0] = r.z, red[1] = the norm left there by k_update_xr / k_init_r (red[2] = ||b||^2 at INIT).
// WITH_DOT = false: plain post-smoothing step (tests, set-up).
template <bool INIT, bool WITH_DOT, int BLOCK>
__global__ void __launch_bounds__(BLOCK)
k_f_post_finish(int64_t n6, const double *__restrict__ b, const double *__restrict__ q, const DinvRef dinv,
                double omega, double *__restrict__ z, double *__restrict__ p, double *partials, unsigned int *counter,
                CgState *state, double *red, int fin_mode, int chk)
{
    ML_RETURN_IF_DONE(state, chk);
    __shared__ __align__(16) double sh[BLOCK];
    static_assert(BLOCK % 6 == 0, "whole nodes per block");
    double rz = 0.0;
    const int64_t stride = (int64_t)gridDim.x * BLOCK;
    const int64_t rounds = (n6 + stride - 1) / stride;
    for (int64_t it = 0; it < rounds; it++) {
        const int64_t t = it * stride + blockIdx.x * (int64_t)BLOCK + threadIdx.x;
        const bool valid = t < n6;
        const double bv = valid ? b[t] : 0.0;
        __syncthreads();
        sh[threadIdx.x] = valid ? bv - q[t] : 0.0;
        __syncthreads();
        if (valid) {
            const double zv = z[t] + omega * dinv_row(dinv, t, sh + (threadIdx.x / 6) * 6);
            z[t] = zv;
            if (INIT) p[t] = zv;
            rz += bv * zv;
        }
    }
    if (!WITH_DOT) return;
    double v[1] = {rz}, out[1];
    if (grid_reduce<1, BLOCK>(v, partials, counter, out) && threadIdx.x == 0) {
        red[0] = out[0];
        if (fin_mode == FIN_INLINE) {
            if (INIT) finalize_init(state, out[0], red[1], red[2]);
            else finalize_update(state, out[0], red[1]);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// lattice levels.  Stencil storage: A[(s*6 + b) * 6n + 6p + a] = entry (a, b) of the block coupling cell p to
// its neighbour in slot s (structure of arrays: thread 6p+a streams its 6*ns values with unit stride).
// ---------------------------------------------------------------------------------------------
enum { LAT_RESID = 0, LAT_RSMOOTH = 1, LAT_PADD = 2, LAT_POST = 3 };

#define LAT_STENCIL_LAUNCH(MODE, geom, grid, stream, ...)                                          \
    do {                                                                                           \
        if ((geom).ns == 9) k_lat_stencil<MODE, 9><<<grid, 192, 0, stream>>>(__VA_ARGS__);         \
        else if ((geom).ns == 27) k_lat_stencil<MODE, 27><<<grid, 192, 0, stream>>>(__VA_ARGS__);  \
        else k_lat_stencil<MODE, 3><<<grid, 192, 0, stream>>>(__VA_ARGS__);                        \
    } while (0)

// v = A in, then
//   RESID    out1 = aux - v (aux null: zero)   out2 = D^+ out1
//   RSMOOTH  out1 = aux - w v                  (out1 may alias aux)
//   PADD     out1 (+)= in - w D^+ v            (flag: accumulate)
//   POST     out1 = in + w D^+ (aux - v)       (out1 must not alias in)
template <int MODE, int NS>
__global__ void __launch_bounds__(192)
k_lat_stencil(const __grid_constant__ LatGeom g, int c0, int c1, const double *__restrict__ A, const double *__restrict__ dinv,
              const double *__restrict__ in, const double *aux, double *out1, double *__restrict__ out2, double omega,
              int flag, const CgState *state, int chk)
{
    ML_RETURN_IF_DONE(state, chk);
    __shared__ __align__(16) double sh[192];
    const int64_t n6 = 6 * (int64_t)g.n;
    const int64_t t = 6 * (int64_t)c0 + blockIdx.x * (int64_t)192 + threadIdx.x;   // cells [c0, c1): this rank's slabs
    const bool valid = t < 6 * (int64_t)c1;
    double v = 0.0;
    if (valid) {
        const int p = (int)(t / 6);
        int k[3];
        lat_unindex(g, p, k);
        // NS is a compile-time constant: the slot loop unrolls, the offsets fold, and all 6*NS value loads are in
        // flight together.  Slots that leave the lattice hold zeros (never written by the probing), so their
        // neighbour is redirected to the cell itself instead of being skipped.
#pragma unroll
        for (int s = 0; s < NS; s++) {
            int o[3];
            lat_stencil_off(g, s, o);
            const int kk[3] = {k[0] + o[0], k[1] + o[1], k[2] + o[2]};
            const bool inside = kk[0] >= 0 && kk[0] < g.np[0] && kk[1] >= 0 && kk[1] < g.np[1] && kk[2] >= 0 && kk[2] < g.np[2];
            const double *xin = in + 6 * (size_t)(inside ? lat_index(g, kk) : p);
            const double *as = A + (size_t)(6 * s) * n6 + t;
#pragma unroll
            for (int b = 0; b < 6; b++) v += as[(size_t)b * n6] * xin[b];
        }
    }
    if (MODE == LAT_RSMOOTH) {
        if (valid) out1[t] = aux[t] - omega * v;
        return;
    }
    double w = v;                                        // what D^+ is applied to
    if (MODE == LAT_RESID || MODE == LAT_POST) w = valid ? (aux ? aux[t] : 0.0) - v : 0.0;
    sh[threadIdx.x] = w;
    __syncthreads();
    if (!valid) return;
    const double dw = dinv_row(dinv, t, sh + (threadIdx.x / 6) * 6);
    if (MODE == LAT_RESID) {
        out1[t] = w;
        out2[t] = dw;
    } else if (MODE == LAT_PADD) {
        const double add = in[t] - omega * dw;
        out1[t] = flag ? out1[t] + add : add;
    } else {
        out1[t] = in[t] + omega * dw;
    }
}

// ---------------------------------------------------------------------------------------------
// lattice blocks of a shell lying in a coordinate plane: the rigid-body modes split into the in-plane class (two
// translations + the rotation about the normal) and the out-of-plane class (deflection + two tilts), which never
// couple -- 18 of 36 entries.  (The mesh blocks have 14: there the drilling rotation couples to nothing, but on a
// lattice the rotation about the normal moves the in-plane translations of the aggregate.)
constexpr unsigned long long LAT_MASK_XY = sell_group(0, 1, 5) | sell_group(2, 3, 4);
constexpr unsigned long long LAT_MASK_XZ = sell_group(0, 2, 4) | sell_group(1, 3, 5);
constexpr unsigned long long LAT_MASK_YZ = sell_group(1, 2, 3) | sell_group(0, 4, 5);
constexpr int LAT_NZ = 18;
static_assert(sell_popcount(LAT_MASK_XY) == LAT_NZ && sell_popcount(LAT_MASK_XZ) == LAT_NZ && sell_popcount(LAT_MASK_YZ) == LAT_NZ, "lattice masks");

// the same four operations on the compacted stencil of a shell lying in a coordinate plane: thread = cell (all six
// rows), 18 value planes per slot instead of 36, every load of a warp one contiguous line.  Skipped entries are exact
// zeros (never written by the paired probing, kept by ginv6), and the sums run in the order of k_lat_stencil /
// dinv_row with those zeros spelled out, so the results are bit-identical to the full-block kernels.
// ---------------------------------------------------------------------------------------------
template <unsigned long long MASK>
__device__ __forceinline__ void dinv_apply_c(const double *__restrict__ Dc, int64_t n, int64_t p, const double w[6], double dw[6])
{
#pragma unroll
    for (int a = 0; a < 6; a++) {
        double s = 0.0;
#pragma unroll
        for (int h = 0; h < 3; h++) {   // dinv_row's pairs; a structural zero is the literal 0.0
            const double dx = (MASK & sell_bit(a, 2 * h)) ? Dc[(size_t)sell_item(MASK, a, 2 * h) * n + p] : 0.0;
            const double dy = (MASK & sell_bit(a, 2 * h + 1)) ? Dc[(size_t)sell_item(MASK, a, 2 * h + 1) * n + p] : 0.0;
            if ((MASK & sell_bit(a, 2 * h)) || (MASK & sell_bit(a, 2 * h + 1))) s += dx * w[2 * h] + dy * w[2 * h + 1];
        }
        dw[a] = s;
    }
}

// rows of mode class CLS: class 0 = the rows coupled to row 0, class 1 = the other three
__host__ __device__ constexpr bool lat_in_class(unsigned long long mask, int cls, int a) { return ((mask & sell_bit(0, a)) != 0) == (cls == 0); }

// thread = (cell, mode class): the two classes of a block never couple, so each is a 3x3 problem of its own with half
// the loads -- twice the threads in flight for a kernel that waits on memory latency (profiles/r02w_ncu_ml.txt)
template <int MODE, int NS, unsigned long long MASK, int CLS>
__device__ __forceinline__ void lat_stencil_c_class(const LatGeom &g, int p, const double *__restrict__ Ac, const double *__restrict__ Dc,
                                                    const double *__restrict__ in, const double *aux, double *out1, double *__restrict__ out2,
                                                    double omega, int flag)
{
    constexpr int NZ = sell_popcount(MASK);
    const int64_t n = g.n;
    int k[3];
    lat_unindex(g, p, k);
    double v[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int s = 0; s < NS; s++) {
        int o[3];
        lat_stencil_off(g, s, o);
        const int kk[3] = {k[0] + o[0], k[1] + o[1], k[2] + o[2]};
        const bool inside = kk[0] >= 0 && kk[0] < g.np[0] && kk[1] >= 0 && kk[1] < g.np[1] && kk[2] >= 0 && kk[2] < g.np[2];
        const double *xin = in + 6 * (size_t)(inside ? lat_index(g, kk) : p);
        double xv[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int b = 0; b < 6; b++)
            if (lat_in_class(MASK, CLS, b)) xv[b] = xin[b];
        const double *as = Ac + (size_t)(NZ * s) * n + p;
#pragma unroll
        for (int a = 0; a < 6; a++)
#pragma unroll
            for (int b = 0; b < 6; b++)
                if (lat_in_class(MASK, CLS, a) && (MASK & sell_bit(a, b))) v[a] += as[(size_t)sell_item(MASK, a, b) * n] * xv[b];
    }
    const size_t at = 6 * (size_t)p;
    if (MODE == LAT_RSMOOTH) {
#pragma unroll
        for (int a = 0; a < 6; a++)
            if (lat_in_class(MASK, CLS, a)) out1[at + a] = aux[at + a] - omega * v[a];
        return;
    }
    double w[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0}, dw[6];   // the other class stays zero: dinv_apply_c multiplies it by literal zeros
#pragma unroll
    for (int a = 0; a < 6; a++)
        if (lat_in_class(MASK, CLS, a)) w[a] = (MODE == LAT_RESID || MODE == LAT_POST) ? (aux ? aux[at + a] : 0.0) - v[a] : v[a];
    dinv_apply_c<MASK>(Dc, n, p, w, dw);
#pragma unroll
    for (int a = 0; a < 6; a++)
        if (lat_in_class(MASK, CLS, a)) {
            if (MODE == LAT_RESID) {
                out1[at + a] = w[a];
                out2[at + a] = dw[a];
            } else if (MODE == LAT_PADD) {
                const double add = in[at + a] - omega * dw[a];
                out1[at + a] = flag ? out1[at + a] + add : add;
            } else {
                out1[at + a] = in[at + a] + omega * dw[a];
            }
        }
}

template <int MODE, int NS, unsigned long long MASK>
__global__ void __launch_bounds__(128)
k_lat_stencil_c(const __grid_constant__ LatGeom g, int c0, int c1, const double *__restrict__ Ac, const double *__restrict__ Dc,
                const double *__restrict__ in, const double *aux, double *out1, double *__restrict__ out2, double omega,
                int flag, const CgState *state, int chk)
{
    ML_RETURN_IF_DONE(state, chk);
    const int p = c0 + blockIdx.x * 128 + threadIdx.x;   // cells [c0, c1): this rank's slabs; blockIdx.y = mode class
    if (p >= c1) return;
    if (blockIdx.y == 0) lat_stencil_c_class<MODE, NS, MASK, 0>(g, p, Ac, Dc, in, aux, out1, out2, omega, flag);
    else lat_stencil_c_class<MODE, NS, MASK, 1>(g, p, Ac, Dc, in, aux, out1, out2, omega, flag);
}

// A / dinv (as probed) -> Ac / Dc for the cells [c0, c1); *bad is set when an entry outside the mask is not an exact zero
template <unsigned long long MASK>
__global__ void k_lat_compact(const __grid_constant__ LatGeom g, int c0, int c1, const double *__restrict__ A, const double *__restrict__ dinv,
                              double *__restrict__ Ac, double *__restrict__ Dc, int *bad)
{
    constexpr int NZ = sell_popcount(MASK);
    const int p = c0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= c1) return;
    const int64_t n = g.n, n6 = 6 * n;
    bool off = false;
    for (int s = 0; s < g.ns; s++)
        for (int a = 0; a < 6; a++)
            for (int b = 0; b < 6; b++) {
                const double val = A[(size_t)(6 * s + b) * n6 + 6 * (size_t)p + a];
                if (MASK & sell_bit(a, b)) Ac[(size_t)(NZ * s + sell_item(MASK, a, b)) * n + p] = val;
                else off |= val != 0.0;
            }
    for (int a = 0; a < 6; a++)
        for (int b = 0; b < 6; b++) {
            const double val = dinv[36 * (size_t)p + 6 * a + b];
            if (MASK & sell_bit(a, b)) Dc[(size_t)sell_item(MASK, a, b) * n + p] = val;
            else off |= val != 0.0;
        }
    if (off) *bad = 1;
}

__global__ void __launch_bounds__(256)
k_lat_smooth0(int64_t t0, int64_t n6, const double *__restrict__ b, const double *__restrict__ dinv, double omega, double *__restrict__ x,
              const CgState *state, int chk)
{
    ML_RETURN_IF_DONE(state, chk);
    const int64_t t = t0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;   // entries [t0, n6)
    if (t >= n6) return;
    x[t] = omega * dinv_row(dinv, t, b + 6 * (t / 6));
}

// parent cell K gathers its (up to 3^d) children: y[K] = sum B_c^T s_c, B_c = rigid-body modes of K at the child centre
// parents [P0, P1); only children in [cc0, cc1) count (a rank's own cells: the sum is partial for a parent whose
// children are split between two ranks)
__global__ void __launch_bounds__(128)
k_lat_restrict(const __grid_constant__ LatGeom gc, const __grid_constant__ LatGeom gp, int P0, int P1, int cc0, int cc1,
               const double *__restrict__ s, double *__restrict__ y, const CgState *state, int chk)
{
    ML_RETURN_IF_DONE(state, chk);
    const int K = P0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (K >= P1) return;
    int kp[3];
    double cp[3], acc[6] = {0, 0, 0, 0, 0, 0};
    lat_unindex(gp, K, kp);
    lat_centre(gp, kp, cp);
    const int n0 = gc.active[0] ? 3 : 1, n1 = gc.active[1] ? 3 : 1, n2 = gc.active[2] ? 3 : 1;
    for (int i2 = 0; i2 < n2; i2++)
        for (int i1 = 0; i1 < n1; i1++)
            for (int i0 = 0; i0 < n0; i0++) {
                const int kc[3] = {gc.active[0] ? 3 * kp[0] + i0 : 0, gc.active[1] ? 3 * kp[1] + i1 : 0, gc.active[2] ? 3 * kp[2] + i2 : 0};
                if (kc[0] >= gc.np[0] || kc[1] >= gc.np[1] || kc[2] >= gc.np[2]) continue;
                const int ci = lat_index(gc, kc);
                if (ci < cc0 || ci >= cc1) continue;
                double cc[3], sv[6];
                lat_centre(gc, kc, cc);
                const double rho[3] = {cc[0] - cp[0], cc[1] - cp[1], cc[2] - cp[2]};
                load6(s + 6 * (size_t)ci, sv);
                rbm_apply_t(rho, sv, acc);
            }
    store6(y + 6 * (size_t)K, acc);
}

// child cell: t = B_c e[parent]
__global__ void __launch_bounds__(256)
k_lat_prolong_t(const __grid_constant__ LatGeom gc, const __grid_constant__ LatGeom gp, int c0, int c1, const double *__restrict__ e,
                double *__restrict__ tv, const CgState *state, int chk)
{
    ML_RETURN_IF_DONE(state, chk);
    const int c = c0 + blockIdx.x * blockDim.x + threadIdx.x;   // children [c0, c1)
    if (c >= c1) return;
    int kc[3], kp[3];
    double cc[3], cp[3], ev[6], v[6];
    lat_unindex(gc, c, kc);
    for (int d = 0; d < 3; d++) kp[d] = gc.active[d] ? kc[d] / 3 : 0;
    lat_centre(gc, kc, cc);
    lat_centre(gp, kp, cp);
    const double rho[3] = {cc[0] - cp[0], cc[1] - cp[1], cc[2] - cp[2]};
    load6(e + 6 * (size_t)lat_index(gp, kp), ev);
    rbm_apply(rho, ev, v);
    store6(tv + 6 * (size_t)c, v);
}

// x = Minv b, one warp per row
__global__ void __launch_bounds__(256)
k_dense_matvec(int n, const double *__restrict__ Minv, const double *__restrict__ b, double *__restrict__ x, const CgState *state, int chk)
{
    ML_RETURN_IF_DONE(state, chk);
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= n) return;
    double s = 0.0;
    for (int j = lane; j < n; j += 32) s += Minv[(size_t)row * n + j] * b[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) x[row] = s;
}

// ---------------------------------------------------------------------------------------------
// set-up kernels
// ---------------------------------------------------------------------------------------------
// probing vector on a lattice: unit mode `mode` on every cell of colour col (index mod 3 per active dimension)
// mode2 >= 0: a second unit mode from the OTHER decoupled class rides along (see ml_probe_pairs)
__global__ void k_lat_set_probe(const __grid_constant__ LatGeom g, int c0, int c1, int c2, int mode, int mode2, double *__restrict__ e)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= g.n) return;
    int k[3];
    lat_unindex(g, a, k);
    const bool hit = (!g.active[0] || k[0] % 3 == c0) && (!g.active[1] || k[1] % 3 == c1) && (!g.active[2] || k[2] % 3 == c2);
    double v[6] = {0, 0, 0, 0, 0, 0};
    if (hit) {
        v[mode] = 1.0;
        if (mode2 >= 0) v[mode2] = 1.0;
    }
    store6(e + 6 * (size_t)a, v);
}

// y = -(P^T A P) e for the probing vector above: column `mode` of the block coupling each cell to its one
// neighbour of colour col
// With a paired probe (mode2 >= 0) row i of the response belongs to the column of the mode in ITS class (bit i of
// class_a set: the class of `mode`); its entry in the other column is a structural zero and stays zero.
__global__ void k_lat_collect(const __grid_constant__ LatGeom g, int cell0, int cell1, int c0, int c1, int c2, int mode, int mode2, unsigned class_a,
                              const double *__restrict__ y, double *__restrict__ A)
{
    const int64_t n6 = 6 * (int64_t)g.n;
    const int64_t t = 6 * (int64_t)cell0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;   // this rank's cells
    if (t >= 6 * (int64_t)cell1) return;
    const int a = (int)(t / 6);
    int k[3], o[3];
    lat_unindex(g, a, k);
    const int col[3] = {c0, c1, c2};
    for (int d = 0; d < 3; d++) {
        o[d] = g.active[d] ? ((col[d] - k[d] % 3 + 1) % 3 + 3) % 3 - 1 : 0;
        const int kk = k[d] + o[d];
        if (kk < 0 || kk >= g.np[d]) return;
    }
    const int s = lat_stencil_slot(g, o);
    const int column = (mode2 < 0 || ((class_a >> (int)(t - 6 * (int64_t)a)) & 1u)) ? mode : mode2;
    A[(size_t)(6 * s + column) * n6 + t] = -y[t];
}

// generalised inverse of a symmetric positive semi-definite 6x6 block: Gauss-Jordan with diagonal pivoting; pivots
// below 1e-12 of the largest diagonal entry (empty cells, aggregates whose free DOFs do not carry a mode) are
// left out and their rows/columns zeroed, i.e. the inverse on the pivoted coordinates
__device__ void ginv6(double M[6][6])
{
    bool used[6] = {false, false, false, false, false, false};
    double dmax = 0.0;
    for (int a = 0; a < 6; a++) dmax = fmax(dmax, M[a][a]);
    const double tol = 1e-12 * dmax;
    for (int step = 0; step < 6; step++) {
        int k = -1;
        double best = tol;
        for (int a = 0; a < 6; a++)
            if (!used[a] && M[a][a] > best) {
                best = M[a][a];
                k = a;
            }
        if (k < 0) break;
        used[k] = true;
        const double ip = 1.0 / M[k][k];
        double col[6], row[6];
        for (int j = 0; j < 6; j++) {
            col[j] = M[j][k];
            row[j] = M[k][j] * ip;
        }
        for (int i = 0; i < 6; i++)
            for (int j = 0; j < 6; j++) {
                if (i == k && j == k) M[i][j] = ip;
                else if (i == k) M[i][j] = row[j];
                else if (j == k) M[i][j] = -col[i] * ip;
                else M[i][j] -= col[i] * row[j];
            }
    }
    for (int i = 0; i < 6; i++)
        for (int j = 0; j < 6; j++)
            if (!used[i] || !used[j]) M[i][j] = 0.0;
}

__global__ void k_lat_extract_dinv(const __grid_constant__ LatGeom g, int cell0, int cell1, const double *__restrict__ A, double *__restrict__ dinv)
{
    const int p = cell0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= cell1) return;
    const int64_t n6 = 6 * (int64_t)g.n;
    const int zero[3] = {0, 0, 0};
    const int sc = lat_stencil_slot(g, zero);
    double D[6][6], S[6][6];
    for (int a = 0; a < 6; a++)
        for (int b = 0; b < 6; b++) D[a][b] = A[(size_t)(6 * sc + b) * n6 + 6 * (size_t)p + a];
    for (int a = 0; a < 6; a++)
        for (int b = 0; b < 6; b++) S[a][b] = 0.5 * (D[a][b] + D[b][a]);
    ginv6(S);
    for (int a = 0; a < 6; a++)
        for (int b = 0; b < 6; b++) dinv[36 * (size_t)p + 6 * a + b] = 0.5 * (S[a][b] + S[b][a]);
}

// dense copy of a lattice stencil (coarsest level), M zeroed beforehand
__global__ void k_lat_to_dense(const __grid_constant__ LatGeom g, const double *__restrict__ A, double *__restrict__ M)
{
    const int64_t n6 = 6 * (int64_t)g.n;
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n6) return;
    int k[3];
    lat_unindex(g, (int)(t / 6), k);
    for (int s = 0; s < g.ns; s++) {
        int o[3];
        lat_stencil_off(g, s, o);
        const int kk[3] = {k[0] + o[0], k[1] + o[1], k[2] + o[2]};
        if (kk[0] < 0 || kk[0] >= g.np[0] || kk[1] < 0 || kk[1] >= g.np[1] || kk[2] < 0 || kk[2] >= g.np[2]) continue;
        const int64_t c0 = 6 * (int64_t)lat_index(g, kk);
        for (int b = 0; b < 6; b++) M[t * n6 + c0 + b] = A[(size_t)(6 * s + b) * n6 + t];
    }
}

// M <- (M + M^T) / 2
__global__ void k_dense_symmetrize(int n, double *M)
{
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx >= (int64_t)n * n) return;
    const int i = (int)(idx / n), j = (int)(idx % n);
    if (i >= j) return;
    const double v = 0.5 * (M[(size_t)i * n + j] + M[(size_t)j * n + i]);
    M[(size_t)i * n + j] = v;
    M[(size_t)j * n + i] = v;
}

// In-place Gauss-Jordan inversion of a symmetric positive semi-definite matrix, one launch per pivot, one thread
// block per row.  Pivots below 1e-12 of the largest diagonal entry (empty cells) are skipped and their rows and
// columns zeroed, which yields the inverse on the complement.  The pivot row is read from a snapshot (rowin)
// taken by the previous launch, so no block reads entries another block is overwriting; the block owning row
// k+1 leaves the snapshot for the next launch in rowout.
__global__ void __launch_bounds__(256)
k_dense_diag_max(int n, const double *__restrict__ M, double *__restrict__ dmax, double *__restrict__ row0)
{
    __shared__ double s_part[8];
    double dm = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        dm = fmax(dm, M[(size_t)i * n + i]);
        row0[i] = M[i];
    }
    for (int o = 16; o > 0; o >>= 1) dm = fmax(dm, __shfl_xor_sync(0xffffffffu, dm, o));
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = dm;
    __syncthreads();
    if (threadIdx.x == 0) {
        double m = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) m = fmax(m, s_part[w]);
        dmax[0] = m;
    }
}

__global__ void __launch_bounds__(256)
k_dense_gj_step(int n, int k, double *__restrict__ M, const double *__restrict__ rowin, double *__restrict__ rowout,
                const double *__restrict__ dmax)
{
    const int i = blockIdx.x;
    double *Mi = M + (size_t)i * n;
    const double piv = rowin[k];
    const double ci = Mi[k];
    __syncthreads();  // every thread holds the old M[i][k] before one of them overwrites it
    if (!(piv > 1e-12 * dmax[0])) {
        if (i == k) {
            for (int j = threadIdx.x; j < n; j += blockDim.x) Mi[j] = 0.0;
        } else if ((int)threadIdx.x == k % (int)blockDim.x) Mi[k] = 0.0;  // the thread that owns column k in the loops below
    } else {
        const double ip = 1.0 / piv;
        if (i == k) {
            for (int j = threadIdx.x; j < n; j += blockDim.x) Mi[j] = (j == k) ? ip : rowin[j] * ip;
        } else {
            const double f = ci * ip;
            for (int j = threadIdx.x; j < n; j += blockDim.x) Mi[j] = (j == k) ? -f : Mi[j] - f * rowin[j];
        }
    }
    if (i == k + 1) {  // each thread re-reads exactly the entries it wrote
        for (int j = threadIdx.x; j < n; j += blockDim.x) rowout[j] = Mi[j];
    }
}

// y += s (the neighbour's partial sums of a boundary slab; two addends, so the order does not matter)
__global__ void k_add_into(int64_t n, const double *__restrict__ s, double *__restrict__ y, const CgState *state, int chk)
{
    ML_RETURN_IF_DONE(state, chk);
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t < n) y[t] += s[t];
}

// deterministic pseudo-random start vector of the power iterations
__global__ void k_fill_hash(int64_t n, uint64_t salt, double *__restrict__ v)
{
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n) return;
    uint64_t h = (uint64_t)t + salt;
    h ^= h >> 33; h *= 0xff51afd7ed558ccdULL; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ULL; h ^= h >> 33;
    v[t] = (double)(h >> 11) * (1.0 / 9007199254740992.0) - 0.5;
}

// out[0] = sum v^2 (deterministic)
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK)
k_norm2(int64_t n, const double *__restrict__ v, double *partials, unsigned int *counter, double *out)
{
    double s = 0.0;
    for (int64_t t = blockIdx.x * (int64_t)BLOCK + threadIdx.x; t < n; t += (int64_t)gridDim.x * BLOCK) s += v[t] * v[t];
    double a[1] = {s}, r[1];
    if (grid_reduce<1, BLOCK>(a, partials, counter, r) && threadIdx.x == 0) out[0] = r[0];
}

__global__ void k_scale_copy(int64_t n, const double *__restrict__ in, double f, double *__restrict__ out)
{
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t < n) out[t] = f * in[t];
}

// largest extent of an element along each axis: non-negative doubles order like their bit patterns, so the maximum is an
// integer atomicMax
__global__ void k_elem_extent(int64_t n_elem, int nen, const int32_t *__restrict__ conn, const double *__restrict__ xyz, unsigned long long *hmax)
{
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    double ext[3] = {0.0, 0.0, 0.0};
    if (e < n_elem) {
        const int32_t *en = conn + (size_t)nen * e;
#pragma unroll
        for (int d = 0; d < 3; d++) {
            double lo = xyz[3 * (size_t)en[0] + d], hi = lo;
            for (int k = 1; k < nen; k++) {
                const double v = xyz[3 * (size_t)en[k] + d];
                lo = fmin(lo, v);
                hi = fmax(hi, v);
            }
            ext[d] = hi - lo;
        }
    }
#pragma unroll
    for (int d = 0; d < 3; d++) {
        unsigned long long b = (unsigned long long)__double_as_longlong(ext[d]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) b = max(b, __shfl_xor_sync(0xffffffffu, b, o));
        if ((threadIdx.x & 31) == 0 && b) atomicMax(hmax + d, b);
    }
}

// c->ml_h = largest element extent per axis over the whole mesh: every element is local to at least one rank
static int ml_element_extents(fs_context *c)
{
    DevBuf<unsigned long long> d_h;
    FS_CUDA(c, d_h.alloc(3));
    FS_CUDA(c, cudaMemsetAsync(d_h.p, 0, 3 * sizeof(unsigned long long), c->stream));
    if (c->n_tri) k_elem_extent<<<nblk(c->n_tri, 256), 256, 0, c->stream>>>(c->n_tri, 3, c->d_tri.p, c->d_xyz.p, d_h.p);
    if (c->n_quad) k_elem_extent<<<nblk(c->n_quad, 256), 256, 0, c->stream>>>(c->n_quad, 4, c->d_quad.p, c->d_xyz.p, d_h.p);
    if (c->world > 1)   // bit patterns of non-negative doubles: the integer maximum is the floating-point maximum
        FS_NCCL_ML(c, nccl().AllReduce(d_h.p, d_h.p, 3, ncclUint64, ncclMax, (ncclComm_t)c->comm, c->stream));
    unsigned long long h[3];
    FS_CUDA(c, cudaMemcpyAsync(h, d_h.p, sizeof h, cudaMemcpyDeviceToHost, c->stream));
    FS_CUDA(c, cudaStreamSynchronize(c->stream));
    for (int d = 0; d < 3; d++) memcpy(&c->ml_h[d], &h[d], sizeof(double));
    return FS_OK;
}

// ---------------------------------------------------------------------------------------------
// host: hierarchy
// ---------------------------------------------------------------------------------------------
static int ml_build_geometry(fs_context *c)
{
    MlHier &m = c->ml;
    m.n_lat = 0;
    {
        int rce = ml_element_extents(c);
        if (rce) return rce;
    }
    double ext[3], maxext = 0.0;
    for (int d = 0; d < 3; d++) {
        ext[d] = c->bbox_hi[d] - c->bbox_lo[d];
        maxext = std::max(maxext, ext[d]);
    }
    LatGeom g = {};
    for (int d = 0; d < 3; d++) g.active[d] = (maxext > 0.0 && ext[d] > 1e-9 * maxext) ? 1 : 0;
    if (!(g.active[0] || g.active[1] || g.active[2])) return fail(c, FS_ERR_STATE, "multilevel preconditioner: mesh has no extent");
    // first lattice: cells three element widths wide, enlarged until the cap on cells holds
    double scale = 1.0;
    for (;;) {
        int64_t tot = 1;
        g.ns = 1;
        for (int d = 0; d < 3; d++) {
            if (g.active[d]) {
                const double h = c->ml_h[d] > 0.0 ? c->ml_h[d] : ext[d];
                g.H[d] = 3.0 * h * scale;
                g.lo[d] = c->bbox_lo[d] - 0.5 * h;
                g.np[d] = (int)std::floor((c->bbox_hi[d] - g.lo[d]) / g.H[d]) + 1;
                g.ns *= 3;
            } else {
                g.H[d] = 0.0;
                g.lo[d] = c->bbox_lo[d];
                g.np[d] = 1;
            }
            tot *= g.np[d];
        }
        if (tot <= std::max<int64_t>(c->ml_max_points, 64)) {
            g.n = (int)tot;
            break;
        }
        scale *= 1.2;
    }
    for (int l = 0; l < ML_MAX_LEVELS; l++) {
        MlLevelBuf &L = m.lat[l];
        L.g = g;
        L.dense = (g.n <= c->ml_dense_points) || (l == ML_MAX_LEVELS - 1);
        m.n_lat = l + 1;
        if (L.dense) break;
        int64_t tot = 1;
        for (int d = 0; d < 3; d++)
            if (g.active[d]) {
                g.np[d] = (g.np[d] + 2) / 3;
                g.H[d] *= 3.0;
                tot *= g.np[d];
            }
        g.n = (int)tot;
    }
    if (m.lat[m.n_lat - 1].g.n > ML_DENSE_MAX_POINTS) return fail(c, FS_ERR_STATE, "multilevel preconditioner: coarsest lattice too large");

    for (int l = 0; l < m.n_lat; l++) {
        MlLevelBuf &L = m.lat[l];
        const size_t n6 = 6 * (size_t)L.g.n;
        FS_CUDA(c, L.A.alloc((size_t)L.g.ns * 6 * n6));
        FS_CUDA(c, L.dinv.alloc(6 * n6));
        FS_CUDA(c, L.x.alloc(n6));
        FS_CUDA(c, L.xb.alloc(n6));
        FS_CUDA(c, L.b.alloc(n6));
        FS_CUDA(c, L.r.alloc(n6));
        FS_CUDA(c, L.t.alloc(n6));
        if (L.dense) FS_CUDA(c, L.minv.alloc(n6 * n6));
        L.omega = 0.0;
    }

    // aggregates of the first lattice: cell of every LOCAL node, and per cell the list of OWNED nodes in it
    const int64_t n_local = c->n_local, n_own = c->n_own, own_lo = c->own_lo;
    std::vector<double> x(3 * (size_t)n_local);
    FS_CUDA(c, cudaMemcpy(x.data(), c->d_xyz.p, sizeof(double) * 3 * n_local, cudaMemcpyDeviceToHost));
    const LatGeom &g1 = m.lat[0].g;
    std::vector<int32_t> agg(n_local), ptr((size_t)g1.n + 1, 0), sup(n_own);
    for (int64_t i = 0; i < n_local; i++) agg[i] = lat_cell_of(g1, &x[3 * i]);
    for (int64_t p = 0; p < n_own; p++) ptr[agg[own_lo + p] + 1]++;
    for (int a = 0; a < g1.n; a++) ptr[a + 1] += ptr[a];
    {
        std::vector<int32_t> fill(ptr.begin(), ptr.end() - 1);
        for (int64_t p = 0; p < n_own; p++) sup[fill[agg[own_lo + p]]++] = (int32_t)p;
    }
    FS_CUDA(c, m.d_agg.alloc(n_local));
    FS_CUDA(c, m.d_sup_ptr.alloc(ptr.size()));
    FS_CUDA(c, m.d_sup_node.alloc(std::max<size_t>(1, sup.size())));
    FS_CUDA(c, cudaMemcpy(m.d_agg.p, agg.data(), sizeof(int32_t) * n_local, cudaMemcpyHostToDevice));
    FS_CUDA(c, cudaMemcpy(m.d_sup_ptr.p, ptr.data(), sizeof(int32_t) * ptr.size(), cudaMemcpyHostToDevice));
    if (!sup.empty()) FS_CUDA(c, cudaMemcpy(m.d_sup_node.p, sup.data(), sizeof(int32_t) * sup.size(), cudaMemcpyHostToDevice));
    FS_CUDA(c, m.d_r1.alloc(6 * (size_t)n_local));
    FS_CUDA(c, m.d_t.alloc(6 * (size_t)n_local));
    FS_CUDA(c, cudaMemset(m.d_r1.p, 0, sizeof(double) * 6 * n_local));
    FS_CUDA(c, cudaMemset(m.d_t.p, 0, sizeof(double) * 6 * n_local));
    FS_CUDA(c, m.d_scalar.alloc(4));

    // ---- distribution of the leading lattice levels over the ranks (header comment) ----
    for (int l = 0; l < m.n_lat; l++) {
        MlLevelBuf &L = m.lat[l];
        L.dist = false;
        L.c0 = 0;
        L.c1 = L.g.n;
        int slow = 0;
        for (int d = 0; d < 3; d++)
            if (L.g.active[d]) slow = d;
        L.n_slabs = L.g.np[slow];
        L.slab_len = L.g.n / L.n_slabs;
        L.s0 = 0;
        L.s1 = L.n_slabs;
        L.halo = 1;
        m.bounds[l].clear();
    }
    m.n_dist = 0;
    if (c->world > 1) {
        const int W = c->world, R = c->rank;
        int64_t min_cells = 32768;
        if (const char *e = getenv("FS_ML_DIST_MIN_CELLS")) min_cells = std::max<int64_t>(1, atoll(e));
        MlLevelBuf &L0 = m.lat[0];
        // slabs touched by this rank's owned / local nodes
        int32_t mine[5] = {0, INT32_MAX, -1, INT32_MAX, -1};   // first owned node's slab, owned min/max, local min/max
        for (int64_t i = 0; i < n_local; i++) {
            const int32_t sl = agg[i] / L0.slab_len;
            mine[3] = std::min(mine[3], sl);
            mine[4] = std::max(mine[4], sl);
            if (i >= own_lo && i < own_lo + n_own) {
                if (i == own_lo) mine[0] = sl;
                mine[1] = std::min(mine[1], sl);
                mine[2] = std::max(mine[2], sl);
            }
        }
        DevBuf<int32_t> d_all;
        FS_CUDA(c, d_all.alloc(5 * (size_t)(W + 1)));
        FS_CUDA(c, cudaMemcpyAsync(d_all.p + 5 * W, mine, sizeof mine, cudaMemcpyHostToDevice, c->stream));
        FS_NCCL_ML(c, nccl().AllGather(d_all.p + 5 * W, d_all.p, 5, ncclInt32, (ncclComm_t)c->comm, c->stream));
        std::vector<int32_t> all(5 * (size_t)W);
        FS_CUDA(c, cudaMemcpyAsync(all.data(), d_all.p, sizeof(int32_t) * 5 * W, cudaMemcpyDeviceToHost, c->stream));
        FS_CUDA(c, cudaStreamSynchronize(c->stream));
        std::vector<int> S(W + 1);
        S[0] = 0;
        S[W] = L0.n_slabs;
        for (int j = 1; j < W; j++) S[j] = all[5 * j];
        bool ok = !L0.dense && L0.g.n >= min_cells;
        int H = 1;
        for (int j = 0; j < W && ok; j++) {
            const int omin = all[5 * j + 1], omax = all[5 * j + 2], lmin = all[5 * j + 3], lmax = all[5 * j + 4];
            if (omax < omin) ok = false;                                    // a rank without owned nodes
            if (omin < S[j] || omax > std::min(S[j + 1], L0.n_slabs - 1)) ok = false;   // owned nodes leave [S_j, S_j+1]
            H = std::max(H, std::max(S[j] - lmin, lmax - (S[j + 1] - 1)));
        }
        if (H > 2) ok = false;
        for (int j = 0; j < W && ok; j++)
            if (S[j + 1] - S[j] < std::max(2, H)) ok = false;
        for (int l = 0; l < m.n_lat && ok; l++) {
            MlLevelBuf &L = m.lat[l];
            if (l > 0) {   // parent bounds: ceil(S / 3)
                if (L.dense || L.g.n < min_cells) break;
                std::vector<int> P(W + 1);
                P[0] = 0;
                P[W] = L.n_slabs;
                bool fine_enough = true;
                for (int j = 1; j < W; j++) P[j] = (S[j] + 2) / 3;
                for (int j = 0; j < W; j++)
                    if (P[j + 1] - P[j] < 2) fine_enough = false;
                if (!fine_enough) break;
                S = P;
            }
            L.dist = true;
            L.halo = l == 0 ? H : 1;
            L.s0 = S[R];
            L.s1 = S[R + 1];
            L.c0 = L.s0 * L.slab_len;
            L.c1 = L.s1 * L.slab_len;
            m.bounds[l] = S;
            FS_CUDA(c, L.stage.alloc(6 * (size_t)L.slab_len));
            m.n_dist = l + 1;
        }
        if (getenv("FS_TIMING") && R == 0) fprintf(stderr, "[fs timing] multilevel: %d of %d lattice levels distributed over %d ranks (halo %d)\n", m.n_dist, m.n_lat, W, H);
    }
    c->ml_geom_ready = true;
    c->ml_values_ready = false;
    return FS_OK;
}

// ---------------------------------------------------------------------------------------------
// exchanges of a distributed level (NCCL point-to-point on the context stream; graph-capturable)
// ---------------------------------------------------------------------------------------------
// the `halo` boundary slabs of the neighbours' values into this rank's copy of vec (and ours into theirs)
static int lat_halo(fs_context *c, int l, double *vec)
{
    MlLevelBuf &L = c->ml.lat[l];
    if (!L.dist) return FS_OK;
    const size_t slab = 6 * (size_t)L.slab_len, h = (size_t)L.halo * slab;
    const int R = c->rank, W = c->world;
    ncclComm_t comm = (ncclComm_t)c->comm;
    FS_NCCL_ML(c, nccl().GroupStart());
    if (R > 0) {
        FS_NCCL_ML(c, nccl().Send(vec + 6 * (size_t)L.c0, h, ncclDouble, R - 1, comm, c->stream));
        FS_NCCL_ML(c, nccl().Recv(vec + 6 * (size_t)L.c0 - h, h, ncclDouble, R - 1, comm, c->stream));
    }
    if (R < W - 1) {
        FS_NCCL_ML(c, nccl().Send(vec + 6 * (size_t)L.c1 - h, h, ncclDouble, R + 1, comm, c->stream));
        FS_NCCL_ML(c, nccl().Recv(vec + 6 * (size_t)L.c1, h, ncclDouble, R + 1, comm, c->stream));
    }
    FS_NCCL_ML(c, nccl().GroupEnd());
    return FS_OK;
}

// partial sums of ONE slab that belongs to a neighbour travel there and are added.  up: the slab after this rank's
// last one goes to rank+1 (restriction from the mesh: owned nodes reach into the next rank's first slab); down: the
// slab before this rank's first one goes to rank-1 (restriction between lattices: a parent whose children are split).
// send_it / recv_it: both sides derive them from the shared bounds.
static int lat_reverse_add(fs_context *c, int l, double *vec, bool up, bool send_it, bool recv_it, int chk)
{
    MlLevelBuf &L = c->ml.lat[l];
    const size_t slab = 6 * (size_t)L.slab_len;
    const int R = c->rank;
    ncclComm_t comm = (ncclComm_t)c->comm;
    if (!send_it && !recv_it) return FS_OK;
    FS_NCCL_ML(c, nccl().GroupStart());
    if (send_it) FS_NCCL_ML(c, nccl().Send(up ? vec + 6 * (size_t)L.c1 : vec + 6 * (size_t)L.c0 - slab, slab, ncclDouble, up ? R + 1 : R - 1, comm, c->stream));
    if (recv_it) FS_NCCL_ML(c, nccl().Recv(L.stage.p, slab, ncclDouble, up ? R - 1 : R + 1, comm, c->stream));
    FS_NCCL_ML(c, nccl().GroupEnd());
    if (recv_it)
        k_add_into<<<nblk((int64_t)slab, 256), 256, 0, c->stream>>>((int64_t)slab, L.stage.p, up ? vec + 6 * (size_t)L.c0 : vec + 6 * (size_t)L.c1 - slab,
                                                                    c->d_state.p, chk);
    return FS_OK;
}

// one stencil operation on this rank's cells of level L, on whichever copy of the stencil the level iterates on
template <int MODE, unsigned long long MASK>
static void lat_stencil_launch_c(MlLevelBuf &L, cudaStream_t st, const double *in, const double *aux, double *out1, double *out2, double omega,
                                 int flag, const CgState *state, int chk)
{
    const dim3 grid(nblk((int64_t)(L.c1 - L.c0), 128), 2);   // y = mode class
    if (L.g.ns == 9) k_lat_stencil_c<MODE, 9, MASK><<<grid, 128, 0, st>>>(L.g, L.c0, L.c1, L.Ac.p, L.Dc.p, in, aux, out1, out2, omega, flag, state, chk);
    else k_lat_stencil_c<MODE, 3, MASK><<<grid, 128, 0, st>>>(L.g, L.c0, L.c1, L.Ac.p, L.Dc.p, in, aux, out1, out2, omega, flag, state, chk);
}

template <int MODE>
static void lat_stencil_launch(fs_context *c, MlLevelBuf &L, const double *in, const double *aux, double *out1, double *out2, double omega, int flag,
                               int chk)
{
    cudaStream_t st = c->stream;
    const CgState *state = c->d_state.p;
    if (L.compact_kind == 0) return lat_stencil_launch_c<MODE, LAT_MASK_XY>(L, st, in, aux, out1, out2, omega, flag, state, chk);
    if (L.compact_kind == 1) return lat_stencil_launch_c<MODE, LAT_MASK_XZ>(L, st, in, aux, out1, out2, omega, flag, state, chk);
    if (L.compact_kind == 2) return lat_stencil_launch_c<MODE, LAT_MASK_YZ>(L, st, in, aux, out1, out2, omega, flag, state, chk);
    const int64_t n6 = 6 * (int64_t)(L.c1 - L.c0);
    LAT_STENCIL_LAUNCH(MODE, L.g, nblk(n6, 192), st, L.g, L.c0, L.c1, L.A.p, L.dinv.p, in, aux, out1, out2, omega, flag, state, chk);
}

// ---------------------------------------------------------------------------------------------
// the building blocks of the cycle (enqueue only)
// ---------------------------------------------------------------------------------------------
static const CgState *st_of(fs_context *c) { return c->d_state.p; }
static DinvRef ml_dinv(const fs_context *c)
{
    return c->ml.dinv_mask ? DinvRef{c->ml.d_dinv_c.p, c->ml.dinv_mask} : DinvRef{c->d_minv.p, 0ull};
}

// mesh level: b_1 = P^T (b - A x).  b and x are LOCAL-layout vectors (b may be null: zero); x's halo is refreshed.
static int fine_restrict_chain(fs_context *c, const double *b, double *x, int chk)
{
    MlHier &m = c->ml;
    cudaStream_t st = c->stream;
    const int64_t o6 = 6 * c->own_lo, n6 = 6 * c->n_own;
    int rc = spmv_once(c, x, c->d_q.p, chk != 0);
    if (rc) return rc;
    k_f_resid<<<nblk(n6, 192), 192, 0, st>>>(n6, b ? b + o6 : nullptr, c->d_q.p + o6, ml_dinv(c), m.d_r1.p + o6, m.d_t.p + o6, st_of(c), chk);
    rc = spmv_once(c, m.d_t.p, c->d_q.p, chk != 0);
    if (rc) return rc;
    MlLevelBuf &L1 = m.lat[0];
    const LatGeom &g1 = L1.g;
    // distributed first lattice: the owned nodes reach the rank's own slabs and the first slab of the next rank
    const int a0 = L1.dist ? L1.c0 : 0, a1 = L1.dist ? std::min(L1.s1 + 1, L1.n_slabs) * L1.slab_len : g1.n;
    k_f_restrict<<<nblk(8 * (int64_t)(a1 - a0), 256), 256, 0, st>>>(g1, a0, a1, m.d_sup_ptr.p, m.d_sup_node.p, c->d_xyz.p + 3 * c->own_lo,
                                                                   c->d_mask.p + c->own_lo, m.d_r1.p + o6, c->d_q.p + o6, m.omega0,
                                                                   L1.b.p, st_of(c), chk);
    if (L1.dist) return lat_reverse_add(c, 0, L1.b.p, true, c->rank < c->world - 1, c->rank > 0, chk);
    if (c->world > 1)
        FS_NCCL_ML(c, nccl().AllReduce(L1.b.p, L1.b.p, 6 * (size_t)g1.n, ncclDouble, ncclSum, (ncclComm_t)c->comm, st));
    return FS_OK;
}

// mesh level: x (+)= P e  (e on the first lattice: replicated, or valid on this rank's slabs -- then its halo slabs,
// which hold the cells of the halo nodes, are fetched first; e_is_global: set-up probes are written everywhere)
static int fine_prolong_chain(fs_context *c, double *e, double *x, bool accumulate, int chk, bool e_is_global = false)
{
    MlHier &m = c->ml;
    cudaStream_t st = c->stream;
    const int64_t o6 = 6 * c->own_lo, n6 = 6 * c->n_own;
    if (!e_is_global) {
        int rc = lat_halo(c, 0, e);
        if (rc) return rc;
    }
    k_f_prolong_t<<<nblk(c->n_local, 256), 256, 0, st>>>(m.lat[0].g, c->n_local, m.d_agg.p, c->d_xyz.p, c->d_mask.p, e, m.d_t.p, st_of(c), chk);
    int rc = spmv_local(c, m.d_t.p, c->d_q.p, chk != 0);  // t is complete on owned and halo nodes: no exchange
    if (rc) return rc;
    k_f_prolong_add<<<nblk(n6, 256), 256, 0, st>>>(n6, m.d_t.p + o6, c->d_q.p + o6, ml_dinv(c), m.omega0, x + o6, accumulate ? 1 : 0, st_of(c), chk);
    return FS_OK;
}

// lattice l: N.b = P^T (b - A x) for the next level N (b may be null: zero)
static int lat_restrict_chain(fs_context *c, int l, const double *b, double *x, int chk)
{
    MlHier &m = c->ml;
    MlLevelBuf &L = m.lat[l], &N = m.lat[l + 1];
    cudaStream_t st = c->stream;
    const int64_t n6 = 6 * (int64_t)(L.c1 - L.c0);
    int rc = lat_halo(c, l, x);
    if (rc) return rc;
    lat_stencil_launch<LAT_RESID>(c, L, x, b, L.r.p, L.t.p, L.omega, 0, chk);
    rc = lat_halo(c, l, L.t.p);
    if (rc) return rc;
    lat_stencil_launch<LAT_RSMOOTH>(c, L, L.t.p, L.r.p, L.r.p, nullptr, L.omega, 0, chk);
    if (!L.dist) {
        k_lat_restrict<<<nblk(N.g.n, 128), 128, 0, st>>>(L.g, N.g, 0, N.g.n, 0, L.g.n, L.r.p, N.b.p, st_of(c), chk);
        return FS_OK;
    }
    // parents with at least one child among this rank's slabs; the sum is partial where the children are split
    const int P0 = (L.s0 / 3) * N.slab_len, P1 = ((L.s1 + 2) / 3) * N.slab_len;
    if (!N.dist) FS_CUDA(c, cudaMemsetAsync(N.b.p, 0, sizeof(double) * 6 * (size_t)N.g.n, st));
    k_lat_restrict<<<nblk(P1 - P0, 128), 128, 0, st>>>(L.g, N.g, P0, P1, L.c0, L.c1, L.r.p, N.b.p, st_of(c), chk);
    if (N.dist) {  // the parent slab split with the rank below belongs to that rank
        const std::vector<int> &S = m.bounds[l];
        const int R = c->rank, W = c->world;
        return lat_reverse_add(c, l + 1, N.b.p, false, R > 0 && S[R] % 3 != 0, R < W - 1 && S[R + 1] % 3 != 0, chk);
    }
    FS_NCCL_ML(c, nccl().AllReduce(N.b.p, N.b.p, 6 * (size_t)N.g.n, ncclDouble, ncclSum, (ncclComm_t)c->comm, st));
    return FS_OK;
}

// lattice l: x (+)= P e, e on the next level (replicated, or valid on this rank's slabs of it; e_is_global: probes)
static int lat_prolong_chain(fs_context *c, int l, double *e, double *x, bool accumulate, int chk, bool e_is_global = false)
{
    MlHier &m = c->ml;
    MlLevelBuf &L = m.lat[l], &N = m.lat[l + 1];
    cudaStream_t st = c->stream;
    const int64_t n6 = 6 * (int64_t)(L.c1 - L.c0);
    if (N.dist && !e_is_global) {
        int rc = lat_halo(c, l + 1, e);
        if (rc) return rc;
    }
    // t = P_t e on this rank's cells and one slab beyond on either side, so that the stencil below needs no exchange
    const int t0 = L.dist ? std::max(L.c0 - L.slab_len, 0) : 0, t1 = L.dist ? std::min(L.c1 + L.slab_len, L.g.n) : L.g.n;
    k_lat_prolong_t<<<nblk(t1 - t0, 256), 256, 0, st>>>(L.g, N.g, t0, t1, e, L.t.p, st_of(c), chk);
    lat_stencil_launch<LAT_PADD>(c, L, L.t.p, nullptr, x, nullptr, L.omega, accumulate ? 1 : 0, chk);
    return FS_OK;
}

// digit l (most significant first) of the cycle index; the last digit serves every deeper level
static int ml_visits(int gamma, int l)
{
    int d[8], n = 0;
    for (int g = gamma; g > 0 && n < 8; g /= 10) d[n++] = g % 10;   // least significant first
    return d[std::max(0, n - 1 - l)];
}

// one cycle on lattice level l for the right-hand side in lat[l].b; *out = where the result lives (valid on this
// rank's cells of a distributed level)
static int lat_cycle(fs_context *c, int l, int chk, double **out)
{
    MlHier &m = c->ml;
    MlLevelBuf &L = m.lat[l];
    cudaStream_t st = c->stream;
    *out = L.xb.p;
    if (L.dense) {
        const int64_t n6 = 6 * (int64_t)L.g.n;
        k_dense_matvec<<<nblk(32 * n6, 256), 256, 0, st>>>((int)n6, L.minv.p, L.b.p, L.xb.p, st_of(c), chk);
        return FS_OK;
    }
    const int64_t t0 = 6 * (int64_t)L.c0, t1 = 6 * (int64_t)L.c1;
    k_lat_smooth0<<<nblk(t1 - t0, 256), 256, 0, st>>>(t0, t1, L.b.p, L.dinv.p, L.omega, L.x.p, st_of(c), chk);
    // a dense level is an exact solve of the Galerkin system: after one visit the restricted residual is zero, a second
    // visit would compute a zero correction (same iteration counts, measured: profiles/r02m_ml_cycle_sweep.json)
    const int n_visits = m.lat[l + 1].dense ? 1 : ml_visits(c->ml_gamma, l);
    for (int gmm = 0; gmm < n_visits; gmm++) {
        int rc = lat_restrict_chain(c, l, L.b.p, L.x.p, chk);
        if (rc) return rc;
        double *e = nullptr;
        if (l == 0 && c->prof.on && gmm < 2) cudaEventRecord(c->prof.ev[7 + 2 * gmm], st);
        rc = lat_cycle(c, l + 1, chk, &e);
        if (rc) return rc;
        if (l == 0 && c->prof.on && gmm < 2) cudaEventRecord(c->prof.ev[8 + 2 * gmm], st);
        rc = lat_prolong_chain(c, l, e, L.x.p, true, chk);
        if (rc) return rc;
    }
    int rc = lat_halo(c, l, L.x.p);
    if (rc) return rc;
    lat_stencil_launch<LAT_POST>(c, L, L.x.p, L.b.p, L.xb.p, nullptr, L.omega, 0, chk);
    return FS_OK;
}

// ---------------------------------------------------------------------------------------------
// set-up of the values: smoother weights, coarse stencils, dense coarsest inverse
// ---------------------------------------------------------------------------------------------
static int read_scalar(fs_context *c, const double *d, int count, double *h, bool all_reduce)
{
    if (all_reduce && c->world > 1)
        FS_NCCL_ML(c, nccl().AllReduce(d, const_cast<double *>(d), count, ncclDouble, ncclSum, (ncclComm_t)c->comm, c->stream));
    FS_CUDA(c, cudaMemcpyAsync(h, d, sizeof(double) * count, cudaMemcpyDeviceToHost, c->stream));
    FS_CUDA(c, cudaStreamSynchronize(c->stream));
    return FS_OK;
}

// the reduction scratch must hold the largest grid of any reducing kernel (+ the spare doubles run_pcg uses)
int ml_ensure_partials(fs_context *c)
{
    const size_t need = (size_t)c->sm_count * 64 * 4 + 64;
    if (c->d_partials.n < need) {
        if (c->cg_graph_exec) { cudaGraphExecDestroy(c->cg_graph_exec); c->cg_graph_exec = nullptr; }
        FS_CUDA(c, c->d_partials.alloc(need));
    }
    return FS_OK;
}

constexpr int ML_POWER_ITS = 20;

// largest eigenvalue of D^-1 A on the mesh level (power iteration), with a safety margin
static int fine_lambda(fs_context *c, double *lam)
{
    MlHier &m = c->ml;
    cudaStream_t st = c->stream;
    const int64_t o6 = 6 * c->own_lo, n6 = 6 * c->n_own;
    const int grid = (int)std::min<int64_t>(nblk(n6, 256), (int64_t)c->sm_count * 4);
    int rc = ml_ensure_partials(c);
    if (rc) return rc;
    FS_CUDA(c, cudaMemsetAsync(c->d_counter.p, 0, sizeof(unsigned int), st));
    double *v = c->d_z.p;  // local layout
    k_fill_hash<<<nblk(n6, 256), 256, 0, st>>>(n6, 0x9e3779b97f4a7c15ULL + (uint64_t)(6 * c->own_begin), v + o6);
    double est = 1.0;
    for (int it = 0; it < ML_POWER_ITS; it++) {
        rc = spmv_once(c, v, c->d_q.p, false);
        if (rc) return rc;
        // t = -D^-1 A v ; r1 = -A v
        k_f_resid<<<nblk(n6, 192), 192, 0, st>>>(n6, nullptr, c->d_q.p + o6, ml_dinv(c), m.d_r1.p + o6, m.d_t.p + o6, st_of(c), 0);
        k_norm2<256><<<grid, 256, 0, st>>>(n6, v + o6, c->d_partials.p, c->d_counter.p, m.d_scalar.p);
        k_norm2<256><<<grid, 256, 0, st>>>(n6, m.d_t.p + o6, c->d_partials.p, c->d_counter.p, m.d_scalar.p + 1);
        double h[2];
        rc = read_scalar(c, m.d_scalar.p, 2, h, true);
        if (rc) return rc;
        if (!(h[0] > 0.0) || !(h[1] > 0.0)) return fail(c, FS_ERR_BREAKDOWN, "multilevel set-up: power iteration collapsed");
        est = std::sqrt(h[1] / h[0]);
        k_scale_copy<<<nblk(n6, 256), 256, 0, st>>>(n6, m.d_t.p + o6, 1.0 / std::sqrt(h[1]), v + o6);
    }
    *lam = 1.1 * est;
    return FS_OK;
}

static int lat_lambda(fs_context *c, int l, double *lam)
{
    MlHier &m = c->ml;
    MlLevelBuf &L = m.lat[l];
    cudaStream_t st = c->stream;
    const int64_t t0 = 6 * (int64_t)L.c0, n6 = 6 * (int64_t)(L.c1 - L.c0);   // this rank's entries
    const int grid = (int)std::min<int64_t>(nblk(n6, 256), (int64_t)c->sm_count * 4);
    int rc = ml_ensure_partials(c);
    if (rc) return rc;
    FS_CUDA(c, cudaMemsetAsync(c->d_counter.p, 0, sizeof(unsigned int), st));
    k_fill_hash<<<nblk(n6, 256), 256, 0, st>>>(n6, 0x51ed270b7f4a7c15ULL + (uint64_t)l + (uint64_t)t0, L.x.p + t0);
    double est = 1.0;
    for (int it = 0; it < ML_POWER_ITS; it++) {
        rc = lat_halo(c, l, L.x.p);
        if (rc) return rc;
        // r = -A x ; t = -D^+ A x
        lat_stencil_launch<LAT_RESID>(c, L, L.x.p, nullptr, L.r.p, L.t.p, 0.0, 0, 0);
        k_norm2<256><<<grid, 256, 0, st>>>(n6, L.x.p + t0, c->d_partials.p, c->d_counter.p, m.d_scalar.p);
        k_norm2<256><<<grid, 256, 0, st>>>(n6, L.t.p + t0, c->d_partials.p, c->d_counter.p, m.d_scalar.p + 1);
        double h[2];
        rc = read_scalar(c, m.d_scalar.p, 2, h, L.dist);
        if (rc) return rc;
        if (!(h[0] > 0.0) || !(h[1] > 0.0)) return fail(c, FS_ERR_BREAKDOWN, "multilevel set-up: power iteration collapsed on a lattice");
        est = std::sqrt(h[1] / h[0]);
        k_scale_copy<<<nblk(n6, 256), 256, 0, st>>>(n6, L.t.p + t0, 1.0 / std::sqrt(h[1]), L.x.p + t0);
    }
    *lam = 1.1 * est;
    return FS_OK;
}

// Probing pairs.  On a shell lying in a coordinate plane the in-plane unknowns (two translations and the rotation
// about the normal) and the out-of-plane ones (deflection and the two tilts) never couple -- that is the block
// pattern the compacted SpMV format was detected from (fs_sell.cuh), the rigid-body modes about cell centres IN
// the plane keep it, and so does every coarse stencil.  One in-plane and one out-of-plane unit mode can then be
// probed with the same vector and separated by the class of the responding row: 3 probes per colour instead of 6,
// with bit-identical stencils.
static void ml_probe_pairs(const fs_context *c, int *n, int a[6], int b[6], unsigned *class_a)
{
    if (!c->sell_active || c->sell_kind < 0 || c->sell_kind > 2) return;
    static const int in_plane[3][3] = {{0, 1, 5}, {0, 2, 4}, {1, 2, 3}}, out_plane[3][3] = {{2, 3, 4}, {1, 3, 5}, {0, 4, 5}};
    const int normal = c->sell_kind == 0 ? 2 : (c->sell_kind == 1 ? 1 : 0);
    if (c->bbox_hi[normal] != c->bbox_lo[normal]) return;  // the lattice centres must lie in the plane of the shell
    *n = 3;
    *class_a = 0;
    for (int k = 0; k < 3; k++) {
        a[k] = in_plane[c->sell_kind][k];
        b[k] = out_plane[c->sell_kind][k];
        *class_a |= 1u << a[k];
    }
}

// Shell in a coordinate plane (the condition of the paired probes): the level iterates on the 18-of-36 copy of its stencil.
// FS_ML_COMPACT=0 keeps the full blocks (lab / tests).
static int lat_compact(fs_context *c, int l, bool planar_pairs)
{
    MlLevelBuf &L = c->ml.lat[l];
    L.compact_kind = -1;
    const char *e = getenv("FS_ML_COMPACT");
    if (!planar_pairs || (e && e[0] == '0') || (L.g.ns != 9 && L.g.ns != 3)) return FS_OK;
    // measured (profiles/r02x_compact_threshold.json): with two threads per cell the compacted kernel also wins on lattices
    // that sit in L2; below a couple of thousand cells a visit is a latency chain either way
    int64_t min_cells = 2048;
    if (const char *m = getenv("FS_ML_COMPACT_MIN_CELLS")) min_cells = std::max<int64_t>(1, atoll(m));
    if ((int64_t)(L.c1 - L.c0) < min_cells) return FS_OK;
    const size_t n = (size_t)L.g.n;
    if (L.Ac.n < (size_t)L.g.ns * LAT_NZ * n) FS_CUDA(c, L.Ac.alloc((size_t)L.g.ns * LAT_NZ * n));
    if (L.Dc.n < LAT_NZ * n) FS_CUDA(c, L.Dc.alloc(LAT_NZ * n));
    FS_CUDA(c, cudaMemsetAsync(c->d_flag.p, 0, sizeof(int), c->stream));
    const unsigned grid = nblk((int64_t)(L.c1 - L.c0), 128);
    if (c->sell_kind == 0) k_lat_compact<LAT_MASK_XY><<<grid, 128, 0, c->stream>>>(L.g, L.c0, L.c1, L.A.p, L.dinv.p, L.Ac.p, L.Dc.p, c->d_flag.p);
    else if (c->sell_kind == 1) k_lat_compact<LAT_MASK_XZ><<<grid, 128, 0, c->stream>>>(L.g, L.c0, L.c1, L.A.p, L.dinv.p, L.Ac.p, L.Dc.p, c->d_flag.p);
    else k_lat_compact<LAT_MASK_YZ><<<grid, 128, 0, c->stream>>>(L.g, L.c0, L.c1, L.A.p, L.dinv.p, L.Ac.p, L.Dc.p, c->d_flag.p);
    int bad = 0;
    FS_CUDA(c, cudaMemcpyAsync(&bad, c->d_flag.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    FS_CUDA(c, cudaStreamSynchronize(c->stream));
    if (!bad) L.compact_kind = c->sell_kind;   // an entry outside the pattern (never seen): stay on the full blocks, on this rank only -- same values
    return FS_OK;
}

// shells in a coordinate plane (same condition as the paired probes): the cycle's mesh-level kernels read a 14-per-node copy
// of the diagonal-block inverses.  FS_ML_COMPACT=0 keeps the full blocks.
static int mesh_dinv_compact(fs_context *c)
{
    MlHier &m = c->ml;
    m.dinv_mask = 0;
    int n_probe = 6, pa[6], pb[6];
    unsigned cls = 0;
    ml_probe_pairs(c, &n_probe, pa, pb, &cls);
    const char *e = getenv("FS_ML_COMPACT");
    if (n_probe != 3 || c->plane_rot || (e && e[0] == '0')) return FS_OK;
    static const unsigned long long masks[3] = {SELL_MASK_XY, SELL_MASK_XZ, SELL_MASK_YZ};
    const unsigned long long mask = masks[c->sell_kind];
    const size_t need = (size_t)sell_popcount(mask) * c->n_own;
    if (m.d_dinv_c.n < need) FS_CUDA(c, m.d_dinv_c.alloc(need));
    FS_CUDA(c, cudaMemsetAsync(c->d_flag.p, 0, sizeof(int), c->stream));
    k_f_dinv_compact<<<nblk(c->n_own, 128), 128, 0, c->stream>>>(c->n_own, mask, c->d_minv.p, m.d_dinv_c.p, c->d_flag.p);
    int bad = 0;
    FS_CUDA(c, cudaMemcpyAsync(&bad, c->d_flag.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    FS_CUDA(c, cudaStreamSynchronize(c->stream));
    if (!bad) m.dinv_mask = mask;
    return FS_OK;
}

int ml_prepare(fs_context *c)
{
    PhaseTimer tmg("ml_prepare");
    if (!c->ml_geom_ready) {
        int rc = ml_build_geometry(c);
        if (rc) return rc;
        if (tmg.on) cudaStreamSynchronize(c->stream);
        tmg.lap("geometry (lattices, aggregates, allocation)");
    }
    if (c->ml_values_ready) return FS_OK;
    if (c->cg_graph_exec) {  // a captured iteration carries the smoother weights of the previous values
        cudaGraphExecDestroy(c->cg_graph_exec);
        c->cg_graph_exec = nullptr;
    }
    MlHier &m = c->ml;
    cudaStream_t st = c->stream;
    cudaEvent_t e0 = c->ev0, e1 = c->ev1;
    FS_CUDA(c, cudaEventRecord(e0, st));
    double lam = 1.0;
    int rc = mesh_dinv_compact(c);
    if (rc) return rc;
    rc = fine_lambda(c, &lam);
    if (rc) return rc;
    m.lambda0 = lam;
    m.omega0 = (4.0 / 3.0) / lam;

    for (int l = 0; l < m.n_lat; l++) {
        MlLevelBuf &L = m.lat[l];
        const LatGeom &g = L.g;
        const int64_t n6 = 6 * (int64_t)g.n;
        // stencil of this level by probing the level below through the cycle's own transfer operators
        FS_CUDA(c, cudaMemsetAsync(L.A.p, 0, sizeof(double) * (size_t)g.ns * 6 * n6, st));
        L.compact_kind = -1;
        const int nc0 = g.active[0] ? 3 : 1, nc1 = g.active[1] ? 3 : 1, nc2 = g.active[2] ? 3 : 1;
        int n_probe = 6, probe_a[6] = {0, 1, 2, 3, 4, 5}, probe_b[6] = {-1, -1, -1, -1, -1, -1};
        unsigned class_a = 0;
        ml_probe_pairs(c, &n_probe, probe_a, probe_b, &class_a);
        for (int c2 = 0; c2 < nc2; c2++)
            for (int c1 = 0; c1 < nc1; c1++)
                for (int c0 = 0; c0 < nc0; c0++)
                    for (int pi = 0; pi < n_probe; pi++) {
                        const int mode = probe_a[pi], mode2 = probe_b[pi];
                        k_lat_set_probe<<<nblk(g.n, 256), 256, 0, st>>>(g, c0, c1, c2, mode, mode2, L.xb.p);
                        if (l == 0) {
                            rc = fine_prolong_chain(c, L.xb.p, c->d_z.p, false, 0, true);
                            if (rc) return rc;
                            rc = fine_restrict_chain(c, nullptr, c->d_z.p, 0);
                            if (rc) return rc;
                        } else {
                            rc = lat_prolong_chain(c, l - 1, L.xb.p, m.lat[l - 1].x.p, false, 0, true);
                            if (rc) return rc;
                            rc = lat_restrict_chain(c, l - 1, nullptr, m.lat[l - 1].x.p, 0);
                            if (rc) return rc;
                        }
                        k_lat_collect<<<nblk(6 * (int64_t)(L.c1 - L.c0), 256), 256, 0, st>>>(g, L.c0, L.c1, c0, c1, c2, mode, mode2, class_a, L.b.p, L.A.p);
                    }
        FS_CUDA(c, cudaGetLastError());
        if (L.dense) {
            FS_CUDA(c, cudaMemsetAsync(L.minv.p, 0, sizeof(double) * (size_t)n6 * n6, st));
            k_lat_to_dense<<<nblk(n6, 128), 128, 0, st>>>(g, L.A.p, L.minv.p);
            k_dense_symmetrize<<<nblk(n6 * n6, 256), 256, 0, st>>>((int)n6, L.minv.p);
            // scratch: level vectors of this (dense) level -- r, t hold the two pivot-row snapshots, x the diagonal maximum
            k_dense_diag_max<<<1, 256, 0, st>>>((int)n6, L.minv.p, L.x.p, L.r.p);
            for (int k = 0; k < (int)n6; k++)
                k_dense_gj_step<<<(unsigned int)n6, 256, 0, st>>>((int)n6, k, L.minv.p, (k & 1) ? L.t.p : L.r.p, (k & 1) ? L.r.p : L.t.p, L.x.p);
            k_dense_symmetrize<<<nblk(n6 * n6, 256), 256, 0, st>>>((int)n6, L.minv.p);
            L.omega = 0.0;
            L.lambda = 0.0;
        } else {
            k_lat_extract_dinv<<<nblk(L.c1 - L.c0, 64), 64, 0, st>>>(g, L.c0, L.c1, L.A.p, L.dinv.p);
            rc = lat_compact(c, l, n_probe == 3);
            if (rc) return rc;
            rc = lat_lambda(c, l, &lam);
            if (rc) return rc;
            L.lambda = lam;
            L.omega = (4.0 / 3.0) / lam;
        }
        FS_CUDA(c, cudaGetLastError());
    }
    FS_CUDA(c, cudaEventRecord(e1, st));
    FS_CUDA(c, cudaStreamSynchronize(st));
    FS_CUDA(c, cudaEventElapsedTime(&m.setup_ms, e0, e1));
    tmg.lap("values (weights, stencils, coarsest inverse)");
    c->ml_values_ready = true;
    return FS_OK;
}

// one application inside the CG iteration (or at its start): r (d_r) -> z (d_z) (+ p at INIT), recurrence advanced.
// red == nullptr: plain application without the dot product (tests).
int ml_enqueue_apply(fs_context *c, bool init, double *red, int fin, int vec_grid)
{
    MlHier &m = c->ml;
    cudaStream_t st = c->stream;
    const int64_t o6 = 6 * c->own_lo, n6 = 6 * c->n_own;
    const int chk = (init || !red) ? 0 : 1;
    k_f_smooth0<<<nblk(n6, 256), 256, 0, st>>>(n6, c->d_r.p + o6, ml_dinv(c), m.omega0, c->d_z.p + o6, st_of(c), chk);
    auto mark = [&](int k) { if (c->prof.on) cudaEventRecord(c->prof.ev[k], st); };
    int rc = fine_restrict_chain(c, c->d_r.p, c->d_z.p, chk);
    if (rc) return rc;
    mark(2);
    double *e = nullptr;
    rc = lat_cycle(c, 0, chk, &e);
    if (rc) return rc;
    mark(3);
    rc = fine_prolong_chain(c, e, c->d_z.p, true, chk);
    if (rc) return rc;
    mark(4);
    rc = spmv_once(c, c->d_z.p, c->d_q.p, chk != 0);
    if (rc) return rc;
    constexpr int PB = 192;
    const int grid = std::max(1, vec_grid);
    if (!red)
        k_f_post_finish<false, false, PB><<<grid, PB, 0, st>>>(n6, c->d_r.p + o6, c->d_q.p + o6, ml_dinv(c), m.omega0, c->d_z.p + o6, nullptr,
                                                                 c->d_partials.p, c->d_counter.p, c->d_state.p, nullptr, fin, 0);
    else if (init)
        k_f_post_finish<true, true, PB><<<grid, PB, 0, st>>>(n6, c->d_r.p + o6, c->d_q.p + o6, ml_dinv(c), m.omega0, c->d_z.p + o6, c->d_p.p + o6,
                                                               c->d_partials.p, c->d_counter.p, c->d_state.p, red, fin, 0);
    else
        k_f_post_finish<false, true, PB><<<grid, PB, 0, st>>>(n6, c->d_r.p + o6, c->d_q.p + o6, ml_dinv(c), m.omega0, c->d_z.p + o6, c->d_p.p + o6,
                                                                c->d_partials.p, c->d_counter.p, c->d_state.p, red, fin, 1);
    return FS_OK;
}

}  // namespace fs

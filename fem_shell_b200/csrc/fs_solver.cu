// fs_solver.cu -- preconditioned conjugate gradients on the block-CSR matrix.
//
// Replaces LinearImplicitSystem::solve -> PetscLinearSolver -> KSPSolve (fs.cpp:138, fsp.cpp:271)
// for the PETSc options the reference passes through (-ksp_type cg, -pc_type none|jacobi|pbjacobi,
// doc/implementation.tex:68-72).  All scalars of the recurrence stay on the device (CgState); the
// host only enqueues batches of iterations and polls one flag, so an iteration costs three kernel
// launches and no synchronisation:
//   k_spmv_dot     q = A p,  partial p.q            -> alpha = rz / p.q
//   k_update       x += alpha p, r -= alpha q, z = M^-1 r, partial r.z and ||r||^2 (or ||z||^2)
//                                                   -> beta, convergence flag
//   k_direction    p = z + beta p
// Reductions are deterministic: per-block partials, the last block to arrive sums them in index
// order.  Multi-GPU: halo exchange of p before k_spmv_dot (ncclSend/Recv), ncclAllReduce of the
// partial sums, then a one-thread finalise kernel.
#include <cub/cub.cuh>

#include <cmath>

#include "fs_context.hpp"
#include "fs_cg_device.cuh"
#include "fs_nccl.hpp"
#include "fs_sell.cuh"

namespace fs {

#define FS_NCCL(ctx, call)                                                                     \
    do {                                                                                       \
        ncclResult_t r__ = (call);                                                             \
        if (r__ != ncclSuccess)                                                                \
            return fs::fail(ctx, FS_ERR_COMM, std::string(#call) + ": " + fs::nccl().GetErrorString(r__)); \
    } while (0)

static inline unsigned int nblk(int64_t n, int bs) { return (unsigned int)((n + bs - 1) / bs); }

// which: 0 init (red = rz, nrm2, bnorm2), 1 after spmv (red = pq), 2 after update (red = rz_new, nrm2)
__global__ void k_finalize(CgState *s, const double *red, int which)
{
    if (which == 0) finalize_init(s, red[0], red[1], red[2]);
    else if (s->done) return;
    else if (which == 1) finalize_pq(s, red[0]);
    else finalize_update(s, red[0], red[1]);
}

// ---------------------------------------------------------------------------------------------
// SpMV: one warp per block row (6 scalar rows of one node).  Lane l owns the double2 at position l
// of every one of the six rows (row length 6*deg doubles = 3*deg double2), so each row is one
// fully coalesced 128-bit load per lane and the x pair is loaded once for all six rows.
// ---------------------------------------------------------------------------------------------
// Occupancy is what this kernel lives on (tools/spmv_lab.cu): capped at 64 registers so that 32 warps
// are resident per SM, it streams at 6.25 TB/s on the 1000x1000-node Quad-4 matrix (0.62 of that at 16 warps).
constexpr int SPMV_BLOCK = 128;
constexpr int SPMV_MIN_BLOCKS = 8;

template <bool WITH_DOT, int BLOCK, bool PEER = false>
__global__ void __launch_bounds__(BLOCK, SPMV_MIN_BLOCKS)
k_spmv(int n_own, int own_lo, const int32_t *__restrict__ nptr, const int32_t *__restrict__ nadj,
       const double *__restrict__ vals, const double *x, double *__restrict__ y_own,
       const double *x_own, double *partials, unsigned int *counter, CgState *state,
       double *red, int fin_mode, PeerWin *pw)
{
    if (WITH_DOT ? state->done : (state && state->done)) return;  // without the dot product: checked only when a state is passed
    if (PEER && !peer_halo_wait(pw)) {  // the neighbours' boundary values of x must have landed (ld_x2: fs_sell.cuh)
        if (blockIdx.x == 0 && threadIdx.x == 0) peer_fail(state);
        return;
    }
    const int lane = threadIdx.x & 31;
    const int warps_per_block = BLOCK / 32;
    const int gw = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    const int nw = gridDim.x * warps_per_block;
    double dot = 0.0;
    for (int p = gw; p < n_own; p += nw) {
        const int b0 = nptr[p], deg = nptr[p + 1] - b0;
        const int L2 = 3 * deg;  // double2 per scalar row
        const double2 *base = reinterpret_cast<const double2 *>(vals + (size_t)36 * b0);
        double acc[6] = {0, 0, 0, 0, 0, 0};
        for (int l = lane; l < L2; l += 32) {
            const int j = l / 3, h = l - 3 * j;
            const int col = nadj[b0 + j];
            const double2 xv = ld_x2<PEER>(reinterpret_cast<const double2 *>(x + 6 * (size_t)col + 2 * h),
                                           PEER && (unsigned)(col - own_lo) >= (unsigned)n_own);
            double2 v[6];
#pragma unroll
            for (int a = 0; a < 6; a++) v[a] = __ldcs(base + (size_t)a * L2 + l);
#pragma unroll
            for (int a = 0; a < 6; a++) acc[a] += v[a].x * xv.x + v[a].y * xv.y;
        }
        // transposing butterfly: 6 sums over 32 lanes in 3+2+1+1+1 exchanges
        double t3[4];
        {
            const bool hi = lane & 16;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                double send = hi ? acc[k] : acc[k + 3];
                double keep = hi ? acc[k + 3] : acc[k];
                t3[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
            }
            t3[3] = 0.0;
        }
        double t2[2];
        {
            const bool hi = lane & 8;
#pragma unroll
            for (int k = 0; k < 2; k++) {
                double send = hi ? t3[k] : t3[k + 2];
                double keep = hi ? t3[k + 2] : t3[k];
                t2[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
            }
        }
        double t1;
        {
            const bool hi = lane & 4;
            double send = hi ? t2[0] : t2[1];
            double keep = hi ? t2[1] : t2[0];
            t1 = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
        t1 += __shfl_xor_sync(0xffffffffu, t1, 2);
        t1 += __shfl_xor_sync(0xffffffffu, t1, 1);
        // lane bits: 16 -> rows {0,1,2} vs {3,4,5}; 8 -> k in {0,1} vs {2,3}; 4 -> k even vs odd
        const int row = ((lane & 16) ? 3 : 0) + ((lane & 8) ? 2 : 0) + ((lane & 4) ? 1 : 0);
        const bool writer = ((lane & 3) == 0) && (((lane >> 2) & 3) != 3);
        if (writer) {
            y_own[6 * (size_t)p + row] = t1;
            if (WITH_DOT) dot += t1 * x_own[6 * (size_t)p + row];
        }
    }
    if (WITH_DOT) {
        double v[1] = {dot}, out[1];
        if (grid_reduce<1, BLOCK>(v, partials, counter, out) && threadIdx.x == 0) finish_dot<1>(out, red, fin_mode, state, pw);
    }
}

// ---------------------------------------------------------------------------------------------
// x += alpha p ; r -= alpha q ; z = M^-1 r ; partial r.z and norm
// PC: 0 none, 1 Jacobi (minv = 1/diag, 6 per node), 2 block Jacobi (minv = 6x6 inverse per node)
// one thread per node (6 dofs = three double2), grid-stride
// ---------------------------------------------------------------------------------------------
template <int PC>
__device__ __forceinline__ void apply_pc_node(const double *__restrict__ minv, int64_t p, const double r[6],
                                              double z[6])
{
    if (PC == 0) {
#pragma unroll
        for (int a = 0; a < 6; a++) z[a] = r[a];
    } else if (PC == 1) {
        const double2 *m = reinterpret_cast<const double2 *>(minv + 6 * p);
#pragma unroll
        for (int h = 0; h < 3; h++) {
            double2 mm = m[h];
            z[2 * h] = mm.x * r[2 * h];
            z[2 * h + 1] = mm.y * r[2 * h + 1];
        }
    } else {
        const double2 *m = reinterpret_cast<const double2 *>(minv + 36 * p);
#pragma unroll
        for (int a = 0; a < 6; a++) {
            double s = 0.0;
#pragma unroll
            for (int h = 0; h < 3; h++) {
                double2 mm = m[3 * a + h];
                s += mm.x * r[2 * h] + mm.y * r[2 * h + 1];
            }
            z[a] = s;
        }
    }
}

template <int PC, int NORM, int BLOCK>
__global__ void __launch_bounds__(BLOCK)
k_update(int64_t n_own, double *__restrict__ x, double *__restrict__ r, const double *__restrict__ p,
         const double *__restrict__ q, double *__restrict__ z, const double *__restrict__ minv,
         double *partials, unsigned int *counter, CgState *state, double *red, int fin_mode, PeerWin *pw)
{
    if (state->done) return;
    double alpha;
    if (pw) {  // p.Ap arrives as one partial per rank in the mailbox: finish the sum here (fs_peer.cuh)
        peer_kern_begin(pw, 1);
        double t[1];
        const bool ok = peer_red_wait<1>(pw, t);
        const bool spd = t[0] > 0.0;
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            if (!ok) peer_fail(state);
            else finalize_pq(state, t[0]);
        }
        if (!ok || !spd) return;
        alpha = state->rz / t[0];
    } else alpha = state->alpha;
    double rz = 0.0, nn = 0.0;
    for (int64_t n = blockIdx.x * (int64_t)BLOCK + threadIdx.x; n < n_own; n += (int64_t)gridDim.x * BLOCK) {
        double xv[6], rv[6], pv[6], qv[6], zv[6];
        load6(x + 6 * n, xv);
        load6(r + 6 * n, rv);
        load6(p + 6 * n, pv);
        load6(q + 6 * n, qv);
#pragma unroll
        for (int a = 0; a < 6; a++) {
            xv[a] += alpha * pv[a];
            rv[a] -= alpha * qv[a];
        }
        apply_pc_node<PC>(minv, n, rv, zv);
        store6(x + 6 * n, xv);
        store6(r + 6 * n, rv);
        store6(z + 6 * n, zv);
#pragma unroll
        for (int a = 0; a < 6; a++) {
            rz += rv[a] * zv[a];
            nn += NORM ? zv[a] * zv[a] : rv[a] * rv[a];
        }
    }
    double v[2] = {rz, nn}, out[2];
    if (grid_reduce<2, BLOCK>(v, partials, counter, out) && threadIdx.x == 0) {
        finish_dot<2>(out, red, fin_mode, state, pw);
        if (pw) peer_kern_end(pw, 1);
    }
}

// x += alpha p ; r -= alpha q ; partial ||r||^2 -- first half of k_update for the multilevel preconditioner
// (z needs the globally restricted r, fs_mlpc.cu).  The norm is left in red[0]; the recurrence advances in
// k_ml_prolong_finish.
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK)
k_update_xr(int64_t n_own, double *__restrict__ x, double *__restrict__ r, const double *__restrict__ p,
            const double *__restrict__ q, double *partials, unsigned int *counter, CgState *state, double *red)
{
    if (state->done) return;
    const double alpha = state->alpha;
    double nn = 0.0;
    for (int64_t n = blockIdx.x * (int64_t)BLOCK + threadIdx.x; n < n_own; n += (int64_t)gridDim.x * BLOCK) {
        double xv[6], rv[6], pv[6], qv[6];
        load6(x + 6 * n, xv);
        load6(r + 6 * n, rv);
        load6(p + 6 * n, pv);
        load6(q + 6 * n, qv);
#pragma unroll
        for (int a = 0; a < 6; a++) {
            xv[a] += alpha * pv[a];
            rv[a] -= alpha * qv[a];
            nn += rv[a] * rv[a];
        }
        store6(x + 6 * n, xv);
        store6(r + 6 * n, rv);
    }
    double v[1] = {nn}, out[1];
    if (grid_reduce<1, BLOCK>(v, partials, counter, out) && threadIdx.x == 0) red[0] = out[0];
}

// r = b - q ; partial ||r||^2 and ||b||^2 into red[0], red[1] (multilevel preconditioner: z follows in fs_mlpc.cu)
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK)
k_init_r(int64_t n_own, const double *__restrict__ b, const double *__restrict__ q, double *__restrict__ r,
         double *partials, unsigned int *counter, double *red)
{
    double nn = 0.0, bb = 0.0;
    for (int64_t n = blockIdx.x * (int64_t)BLOCK + threadIdx.x; n < n_own; n += (int64_t)gridDim.x * BLOCK) {
        double bv[6], qv[6], rv[6];
        load6(b + 6 * n, bv);
        load6(q + 6 * n, qv);
#pragma unroll
        for (int a = 0; a < 6; a++) {
            rv[a] = bv[a] - qv[a];
            nn += rv[a] * rv[a];
            bb += bv[a] * bv[a];
        }
        store6(r + 6 * n, rv);
    }
    double v[2] = {nn, bb}, out[2];
    if (grid_reduce<2, BLOCK>(v, partials, counter, out) && threadIdx.x == 0) {
        red[0] = out[0];
        red[1] = out[1];
    }
}

// p = z + beta p
// PushArgs (peer mode with the halo push folded in): the send list of fs_peer.cuh; own_lo6 = offset of the owned
// part inside the local vectors (z and p below point at the owned part)
struct PushArgs {
    int64_t n_send;
    const int32_t *idx, *push_peer, *push_dst;
    unsigned int *push_counter;
    int n_push_blocks;
    int64_t own_lo;
    const uint8_t *is_send;   // per owned node: 1 = on the send list (its new value is written by the pushing thread)
};

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK)
k_direction(int64_t n_own, const double *__restrict__ z, double *__restrict__ p, CgState *state, PeerWin *pw,
            unsigned int *counter, const __grid_constant__ PushArgs push)
{
    // after convergence x is final; p is not needed any more
    if (state->done) return;
    double beta, t[2] = {0.0, 0.0};
    if (pw) {  // r.z and the norm arrive as per-rank partials: finish the sums, beta from the OLD r.z
        peer_kern_begin(pw, 2);
        if (!peer_red_wait<2>(pw, t)) {
            if (blockIdx.x == 0 && threadIdx.x == 0) peer_fail(state);
            return;
        }
        beta = t[0] / state->rz;
    } else beta = state->beta;
    const int n_push = pw ? push.n_push_blocks : 0;   // leading blocks: the neighbours' halo values of the NEW p and nothing else
    if ((int)blockIdx.x < n_push) {
        // The pushing thread forms the value and stores it here and there (the bulk blocks skip send-list nodes: a
        // bulk thread could otherwise update p before the pushing thread has read the old value).  These blocks only
        // push: the system-scope fence before the stamp waits for the NVLink writes, which must not delay bulk work.
        peer_push_inline(pw, push.n_send, push.idx, push.push_peer, push.push_dst, push.push_counter, n_push,
                         [&](int32_t node, int h) {
                             const size_t at = 3 * (size_t)(node - push.own_lo) + h;   // send-list nodes are owned nodes
                             const double2 zz = reinterpret_cast<const double2 *>(z)[at];
                             double2 *pl = reinterpret_cast<double2 *>(p) + at;
                             const double2 pp = *pl;
                             const double2 v = make_double2(zz.x + beta * pp.x, zz.y + beta * pp.y);
                             *pl = v;
                             return v;
                         });
    } else {
        const int64_t n2 = 3 * n_own;
        const double2 *z2 = reinterpret_cast<const double2 *>(z);
        double2 *p2 = reinterpret_cast<double2 *>(p);
        const uint8_t *skip = n_push > 0 ? push.is_send : nullptr;
        const int64_t stride = (int64_t)(gridDim.x - n_push) * BLOCK;
        for (int64_t i = (blockIdx.x - n_push) * (int64_t)BLOCK + threadIdx.x; i < n2; i += stride) {
            const uint8_t sk = skip ? skip[i / 3] : (uint8_t)0;   // issued together with the two loads below
            double2 zz = z2[i], pp = p2[i];
            pp.x = zz.x + beta * pp.x;
            pp.y = zz.y + beta * pp.y;
            if (!sk) p2[i] = pp;
        }
    }
    if (pw) {  // the block that finishes last advances the recurrence (every block has read the old state)
        __syncthreads();
        if (threadIdx.x == 0 && atomicInc(counter, gridDim.x - 1) == gridDim.x - 1) {
            finalize_update(state, t[0], t[1]);
            peer_kern_end(pw, 2);
        }
    }
}

// r = b - q ; z = M^-1 r ; p = z ; sums r.z, norm(r|z), norm(b|M^-1 b)
template <int PC, int NORM, int BLOCK>
__global__ void __launch_bounds__(BLOCK)
k_init(int64_t n_own, const double *__restrict__ b, const double *__restrict__ q, double *__restrict__ r,
       double *__restrict__ z, double *__restrict__ p, const double *__restrict__ minv, double *partials,
       unsigned int *counter, CgState *state, double *red, int fin_mode, PeerWin *pw)
{
    double rz = 0.0, nn = 0.0, bb = 0.0;
    for (int64_t n = blockIdx.x * (int64_t)BLOCK + threadIdx.x; n < n_own; n += (int64_t)gridDim.x * BLOCK) {
        double bv[6], qv[6], rv[6], zv[6], mb[6];
        load6(b + 6 * n, bv);
        load6(q + 6 * n, qv);
#pragma unroll
        for (int a = 0; a < 6; a++) rv[a] = bv[a] - qv[a];
        apply_pc_node<PC>(minv, n, rv, zv);
        store6(r + 6 * n, rv);
        store6(z + 6 * n, zv);
        store6(p + 6 * n, zv);
        if (NORM) apply_pc_node<PC>(minv, n, bv, mb);
#pragma unroll
        for (int a = 0; a < 6; a++) {
            rz += rv[a] * zv[a];
            nn += NORM ? zv[a] * zv[a] : rv[a] * rv[a];
            bb += NORM ? mb[a] * mb[a] : bv[a] * bv[a];
        }
    }
    double v[3] = {rz, nn, bb}, out[3];
    if (grid_reduce<3, BLOCK>(v, partials, counter, out) && threadIdx.x == 0) finish_dot<3>(out, red, fin_mode, state, pw);
}

// peer mode: the three sums of k_init, one partial per rank -> start of the recurrence
__global__ void k_finalize_init_peer(CgState *s, PeerWin *pw)
{
    double t[3];
    const bool ok = peer_red_wait<3>(pw, t);
    if (threadIdx.x == 0) {
        finalize_init(s, t[0], t[1], t[2]);
        if (!ok) peer_fail(s);
    }
}

// ---------------------------------------------------------------------------------------------
// preconditioner setup: diagonal (or 6x6 diagonal block inverse) of the assembled matrix
// ---------------------------------------------------------------------------------------------
__global__ void k_extract_minv(int n_own, int own_lo, const int32_t *__restrict__ nptr,
                               const int32_t *__restrict__ nadj, const double *__restrict__ vals, int pc,
                               double *minv, int *bad)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_own) return;
    const int b0 = nptr[p], deg = nptr[p + 1] - b0;
    int slot = -1;
    for (int j = 0; j < deg; j++)
        if (nadj[b0 + j] == p + own_lo) slot = j;
    if (slot < 0) { *bad = 1; return; }
    const double *blk = vals + (size_t)36 * b0 + 6 * slot;
    const int L = 6 * deg;
    if (pc == 1) {  // PCJacobi: a zero diagonal entry is replaced by 1 (PETSc's PCSetUp_Jacobi does the same)
        for (int a = 0; a < 6; a++) {
            const double d = blk[(size_t)a * L + a];
            minv[6 * (size_t)p + a] = d != 0.0 ? 1.0 / d : 1.0;
        }
        return;
    }
    double M[6][12];
    for (int a = 0; a < 6; a++)
        for (int b = 0; b < 6; b++) {
            M[a][b] = blk[(size_t)a * L + b];
            M[a][6 + b] = (a == b) ? 1.0 : 0.0;
        }
    for (int c = 0; c < 6; c++) {
        int piv = c;
        for (int r = c + 1; r < 6; r++)
            if (fabs(M[r][c]) > fabs(M[piv][c])) piv = r;
        if (M[piv][c] == 0.0) { *bad = 1; return; }
        if (piv != c)
            for (int j = 0; j < 12; j++) { double t = M[c][j]; M[c][j] = M[piv][j]; M[piv][j] = t; }
        double d = 1.0 / M[c][c];
        for (int j = 0; j < 12; j++) M[c][j] *= d;
        for (int r = 0; r < 6; r++)
            if (r != c) {
                double f = M[r][c];
                for (int j = 0; j < 12; j++) M[r][j] -= f * M[c][j];
            }
    }
    for (int a = 0; a < 6; a++)
        for (int b = 0; b < 6; b++) minv[36 * (size_t)p + 6 * a + b] = M[a][6 + b];
}

int solver_query_occupancy(fs_context *c)
{
    int nb = 0;
    FS_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_spmv<true, SPMV_BLOCK>, SPMV_BLOCK, 0));
    c->spmv_blocks_per_sm = std::max(1, nb);
    FS_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_update<1, 0, 256>, 256, 0));
    c->vec_blocks_per_sm = std::max(1, nb);
    FS_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_spmv_sell<SELL_MASK_XY, true, SELL_BLOCK, SELL_MINB, false, false>, SELL_BLOCK, 0));
    c->sell_blocks_per_sm = std::max(1, nb);
    return FS_OK;
}

// ---------------------------------------------------------------------------------------------
// zero-compacted SpMV copy (fs_sell.cuh): measured and rebuilt lazily after every values pass
// ---------------------------------------------------------------------------------------------
static const unsigned long long SELL_MASKS[3] = {SELL_MASK_XY, SELL_MASK_XZ, SELL_MASK_YZ};

static void drop_cg_graph(fs_context *c)
{
    if (c->cg_graph_exec) cudaGraphExecDestroy(c->cg_graph_exec);
    c->cg_graph_exec = nullptr;
}

// shared memory of k_sell_fill_t for the widest slice of this mesh, or 0 when a slice does not fit an SM
static size_t sell_fill_smem(const fs_context *c)
{
    const size_t bytes = (size_t)32 * ((36 * (size_t)c->sell_dmax_max) | 1) * sizeof(double);
    return bytes <= 200 * 1024 ? bytes : 0;
}

static int sell_fill(fs_context *c, unsigned long long mask, int nz, int write_adj)
{
    const int n_own = (int)c->n_own, n_slices = (int)c->sell_slices;
    const size_t smem = sell_fill_smem(c);
    if (smem) {
        FS_CUDA(c, cudaFuncSetAttribute(k_sell_fill_t, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_sell_fill_t<<<n_slices, SELL_FILL_THREADS, smem, c->stream>>>(n_own, (int)c->own_lo, mask, c->d_nptr.p, c->d_nadj.p, c->d_vals.p,
                                                                        c->d_sell_sptr.p, c->d_sell_adj.p, c->d_sell_vals.p, nz, write_adj,
                                                                        c->d_sell_mask.p);
    } else {
        k_sell_fill<<<nblk(n_slices, 8), 256, 0, c->stream>>>(n_own, n_slices, (int)c->own_lo, mask, c->d_nptr.p, c->d_nadj.p, c->d_vals.p,
                                                               c->d_sell_sptr.p, c->d_sell_adj.p, c->d_sell_vals.p, nz, write_adj);
    }
    FS_CUDA(c, cudaGetLastError());
    return FS_OK;
}

int spmv_format_prepare(fs_context *c)
{
    if (c->sell_checked) return FS_OK;
    if (!c->assembled) return fail(c, FS_ERR_STATE, "matrix not assembled");
    cudaStream_t st = c->stream;
    const bool was = c->sell_active;
    const unsigned long long old_mask = c->sell_mask;
    c->sell_active = false;
    {   // everything below reads the parity values; the slice pass (fs_slice_asm.cu) does not write them
        int rcp = ensure_parity_values(c);
        if (rcp) return rcp;
    }
    c->sell_checked = true;
    if (c->spmv_format_pref == FS_SPMV_FULL) {
        if (was) drop_cg_graph(c);
        return FS_OK;
    }
    const int n_own = (int)c->n_own;
    if (c->d_sell_mask.n < 1) FS_CUDA(c, c->d_sell_mask.alloc(1));
    FS_CUDA(c, cudaMemsetAsync(c->d_sell_mask.p, 0, sizeof(unsigned long long), st));
    unsigned long long m = 0;
    // Re-assembly of a mesh that iterated on the compacted copy before: copy with the same mask in one pass and
    // let the copy kernel report any entry outside it; only then fall back to detect -> decide -> fill.
    if (was && c->sell_layout_ready && sell_fill_smem(c) && c->d_sell_vals.n >= (size_t)32 * c->sell_nz * c->sell_slots) {
        int rc = sell_fill(c, old_mask, c->sell_nz, 0);
        if (rc) return rc;
        FS_CUDA(c, cudaMemcpyAsync(&m, c->d_sell_mask.p, sizeof m, cudaMemcpyDeviceToHost, st));
        FS_CUDA(c, cudaStreamSynchronize(st));
        if ((m & ~old_mask) == 0) {
            c->sell_active = true;
            return FS_OK;
        }
        FS_CUDA(c, cudaMemsetAsync(c->d_sell_mask.p, 0, sizeof(unsigned long long), st));
    }
    k_sell_detect<<<c->sm_count * 8, 256, 0, st>>>(n_own, c->d_nptr.p, c->d_vals.p, c->d_sell_mask.p);
    FS_CUDA(c, cudaMemcpyAsync(&m, c->d_sell_mask.p, sizeof m, cudaMemcpyDeviceToHost, st));
    FS_CUDA(c, cudaStreamSynchronize(st));
    c->sell_detected = m;
    unsigned long long mask = 0;
    for (int k = 0; k < 3 && !mask; k++)
        if ((m & ~SELL_MASKS[k]) == 0) { mask = SELL_MASKS[k]; c->sell_kind = k; }
    if (!mask) {  // blocks are (nearly) dense: the parity format is the SpMV format
        if (was) drop_cg_graph(c);
        return FS_OK;
    }
    {
        int rcl = sell_layout_build(c);   // slice pointers + column nodes, once per mesh
        if (rcl) return rcl;
    }
    const int write_adj = 0;
    const int nz = sell_popcount(mask);
    const size_t need = (size_t)32 * nz * c->sell_slots;
    if (c->d_sell_vals.n < need) FS_CUDA(c, c->d_sell_vals.alloc(need));
    FS_CUDA(c, cudaMemsetAsync(c->d_sell_mask.p, 0, sizeof(unsigned long long), st));
    int rc = sell_fill(c, mask, nz, write_adj);
    if (rc) return rc;
    c->sell_active = true;
    c->sell_mask = mask;
    c->sell_nz = nz;
    if (!was || old_mask != mask) drop_cg_graph(c);
    return FS_OK;
}

template <unsigned long long MASK, bool WITH_DOT, bool PEER, bool ROT>
static void launch_sell_k(fs_context *c, int grid, const double *x, double *y_own, const double *x_own, double *red, int fin_mode, PeerWin *pw,
                          CgState *state)
{
    PlaneQ q;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) q.m[i][j] = c->plane_Q[3 * i + j];
    k_spmv_sell<MASK, WITH_DOT, SELL_BLOCK, SELL_MINB, PEER, ROT><<<grid, SELL_BLOCK, 0, c->stream>>>(
        (int)c->n_own, (int)c->own_lo, (int)c->sell_slices, c->d_sell_sptr.p, c->d_sell_adj.p, c->d_sell_vals.p, x, y_own, x_own,
        c->d_partials.p, c->d_counter.p, state, red, fin_mode, pw, PEER ? c->d_sell_hflag.p : nullptr, q);
}

template <unsigned long long MASK, bool WITH_DOT>
static void launch_sell(fs_context *c, const double *x, double *y_own, const double *x_own, double *red, int fin_mode, PeerWin *pw,
                        CgState *state)
{
    const int64_t want = (c->sell_slices + SELL_BLOCK / 32 - 1) / (SELL_BLOCK / 32);
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)c->sm_count * c->sell_blocks_per_sm));
    // ROT (plane frame of a shell in general position) only exists for the xy pattern: the slice pass produces no other
    const bool rot = c->plane_rot && MASK == SELL_MASK_XY && !c->parity_valid;
    if (WITH_DOT && pw) {
        if (rot) launch_sell_k<SELL_MASK_XY, WITH_DOT, WITH_DOT, true>(c, grid, x, y_own, x_own, red, fin_mode, pw, state);
        else launch_sell_k<MASK, WITH_DOT, WITH_DOT, false>(c, grid, x, y_own, x_own, red, fin_mode, pw, state);
    } else {
        if (rot) launch_sell_k<SELL_MASK_XY, WITH_DOT, false, true>(c, grid, x, y_own, x_own, red, fin_mode, pw, state);
        else launch_sell_k<MASK, WITH_DOT, false, false>(c, grid, x, y_own, x_own, red, fin_mode, pw, state);
    }
}

int solver_prepare(fs_context *c, int pc)
{
    if (!c->assembled) return fail(c, FS_ERR_STATE, "fs_solve before fs_assemble");
    if (pc < 0 || pc > 3) return fail(c, FS_ERR_ARG, "unknown preconditioner");
    {
        int rc = spmv_format_prepare(c);
        if (rc) return rc;
    }
    if (pc == FS_PC_MLRBM) {  // multigrid cycle whose mesh-level smoother is block Jacobi (fs_mlpc.cu)
        int rc = solver_prepare(c, FS_PC_BJACOBI6);
        if (rc) return rc;
        rc = ml_ensure_partials(c);
        if (rc) return rc;
        return ml_prepare(c);
    }
    if (c->minv_kind == pc) return FS_OK;
    if (pc != FS_PC_NONE) {
        const size_t need = (size_t)(pc == 1 ? 6 : 36) * c->n_own;
        if (c->d_minv.n < need) {  // the captured CG graph holds this pointer: only re-capture when it moves
            if (c->cg_graph_exec) { cudaGraphExecDestroy(c->cg_graph_exec); c->cg_graph_exec = nullptr; }
            FS_CUDA(c, c->d_minv.alloc(need));
        }
        DevBuf<int> &bad = c->d_flag;
        FS_CUDA(c, cudaMemsetAsync(bad.p, 0, sizeof(int), c->stream));
        if (c->sell_active && c->sell_mask == SELL_MASK_XY && !c->parity_valid) {  // the slice pass wrote the compacted format only
            int rcm = extract_minv_sell(c, pc, bad.p);
            if (rcm) return rcm;
        } else {
            int rcp = ensure_parity_values(c);
            if (rcp) return rcp;
            k_extract_minv<<<nblk(c->n_own, 128), 128, 0, c->stream>>>((int)c->n_own, (int)c->own_lo, c->d_nptr.p,
                                                                        c->d_nadj.p, c->d_vals.p, pc, c->d_minv.p, bad.p);
        }
        int h_bad = 0;
        FS_CUDA(c, cudaMemcpyAsync(&h_bad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        FS_CUDA(c, cudaStreamSynchronize(c->stream));
        if (h_bad) return fail(c, FS_ERR_BREAKDOWN, "singular diagonal block in the preconditioner");
    }
    c->minv_kind = pc;
    return FS_OK;
}

// ---------------------------------------------------------------------------------------------
// halo exchange (multi-GPU): owned boundary values -> neighbours' halo segments
// ---------------------------------------------------------------------------------------------
__global__ void k_pack(int64_t n_send, const int32_t *__restrict__ idx, const double *__restrict__ vec,
                       double *__restrict__ buf)
{
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= 3 * n_send) return;
    int64_t s = t / 3;
    int h = (int)(t - 3 * s);
    reinterpret_cast<double2 *>(buf)[t] = reinterpret_cast<const double2 *>(vec + 6 * (size_t)idx[s])[h];
}

int halo_exchange(fs_context *c, double *d_vec)
{
    if (c->world == 1 || c->peers.empty()) return FS_OK;
    if (c->send_total)
        k_pack<<<nblk(3 * c->send_total, 256), 256, 0, c->stream>>>(c->send_total, c->d_send_idx.p, d_vec, c->d_sendbuf.p);
    FS_NCCL(c, nccl().GroupStart());
    for (const Peer &pr : c->peers) {
        if (pr.send_count)
            FS_NCCL(c, nccl().Send(c->d_sendbuf.p + 6 * pr.send_off, 6 * pr.send_count, ncclDouble, pr.rank, (ncclComm_t)c->comm, c->stream));
        if (pr.recv_count)
            FS_NCCL(c, nccl().Recv(d_vec + 6 * pr.recv_off, 6 * pr.recv_count, ncclDouble, pr.rank, (ncclComm_t)c->comm, c->stream));
    }
    FS_NCCL(c, nccl().GroupEnd());
    return FS_OK;
}

static int spmv_grid(fs_context *c)
{
    // persistent grid: a multiple of the SM count, one warp per block row, grid-stride
    int64_t want = ((int64_t)c->n_own + SPMV_BLOCK / 32 - 1) / (SPMV_BLOCK / 32);
    int64_t cap = (int64_t)c->sm_count * c->spmv_blocks_per_sm;
    return (int)std::max<int64_t>(1, std::min(want, cap));
}

static int vec_grid(fs_context *c)
{
    int64_t want = ((int64_t)c->n_own + 255) / 256;
    int64_t cap = (int64_t)c->sm_count * c->vec_blocks_per_sm;
    return (int)std::max<int64_t>(1, std::min(want, cap));
}

// y_own = A x (+ partial x_own.y_own) on whichever copy of the matrix is current
template <bool WITH_DOT>
static void launch_spmv(fs_context *c, const double *x, double *y_own, const double *x_own, double *red, int fin_mode, PeerWin *pw,
                        bool check_done = true)
{
    // the kernels with the dot product always honour the done flag; the plain ones only when handed the state
    CgState *state = (WITH_DOT || check_done) ? c->d_state.p : nullptr;
    if (c->sell_active) {
        if (c->sell_mask == SELL_MASK_XY) launch_sell<SELL_MASK_XY, WITH_DOT>(c, x, y_own, x_own, red, fin_mode, pw, state);
        else if (c->sell_mask == SELL_MASK_XZ) launch_sell<SELL_MASK_XZ, WITH_DOT>(c, x, y_own, x_own, red, fin_mode, pw, state);
        else launch_sell<SELL_MASK_YZ, WITH_DOT>(c, x, y_own, x_own, red, fin_mode, pw, state);
        return;
    }
    if (WITH_DOT && pw)
        k_spmv<WITH_DOT, SPMV_BLOCK, WITH_DOT><<<spmv_grid(c), SPMV_BLOCK, 0, c->stream>>>(
            (int)c->n_own, (int)c->own_lo, c->d_nptr.p, c->d_nadj.p, c->d_vals.p, x, y_own, x_own, c->d_partials.p, c->d_counter.p,
            state, red, fin_mode, pw);
    else
        k_spmv<WITH_DOT, SPMV_BLOCK, false><<<spmv_grid(c), SPMV_BLOCK, 0, c->stream>>>(
            (int)c->n_own, (int)c->own_lo, c->d_nptr.p, c->d_nadj.p, c->d_vals.p, x, y_own, x_own, c->d_partials.p, c->d_counter.p,
            state, red, fin_mode, pw);
}

int spmv_once(fs_context *c, const double *d_in, double *d_out, bool check_done)
{
    int rc = spmv_format_prepare(c);
    if (rc) return rc;
    rc = halo_exchange(c, const_cast<double *>(d_in));
    if (rc) return rc;
    launch_spmv<false>(c, d_in, d_out + 6 * c->own_lo, nullptr, nullptr, FIN_RED, nullptr, check_done);
    FS_CUDA(c, cudaGetLastError());
    return FS_OK;
}

int spmv_local(fs_context *c, const double *d_in, double *d_out, bool check_done)
{
    int rc = spmv_format_prepare(c);
    if (rc) return rc;
    launch_spmv<false>(c, d_in, d_out + 6 * c->own_lo, nullptr, nullptr, FIN_RED, nullptr, check_done);
    FS_CUDA(c, cudaGetLastError());
    return FS_OK;
}

template <int PC, int NORM>
static int enqueue_iteration(fs_context *c, double *red, int sg, int vg)
{
    const bool single = (c->world == 1), peer = !single && c->peer_ready && PC != 3;
    const int fin = single ? FIN_INLINE : (peer ? FIN_PEER : FIN_RED);
    PeerWin *pw = peer ? c->d_pw.p : nullptr;
    const int64_t o6 = 6 * c->own_lo;
    auto mark = [&](int k) { if (PC == 3 && c->prof.on) cudaEventRecord(c->prof.ev[k], c->stream); };
    mark(0);
    PushArgs push = {};
    if (peer) {  // the halo of p was pushed by the kernel that formed it (k_direction below, k_halo_push before the first iteration)
        push.n_send = c->send_total;
        push.idx = c->d_send_idx.p;
        push.push_peer = c->d_push_peer.p;
        push.push_dst = c->d_push_dst.p;
        push.push_counter = c->d_counter.p + 1;
        push.n_push_blocks = (int)std::min<int64_t>(vg, std::max<int64_t>(1, nblk(3 * c->send_total, 256)));
        push.own_lo = c->own_lo;
        push.is_send = c->d_is_send.p;
        // send list larger than the grid, or a node sent to two neighbours (two threads would update it), on ANY rank
        // (decided collectively in peer_window_setup): separate kernel at the start of the iteration
        if (!c->push_foldable) push.n_push_blocks = 0;
        if (push.n_push_blocks == 0)
            k_halo_push<<<std::max(1u, nblk(3 * c->send_total, 256)), 256, 0, c->stream>>>(
                pw, c->send_total, c->d_send_idx.p, c->d_push_peer.p, c->d_push_dst.p, c->d_p.p, c->d_state.p, c->d_counter.p + 1);
    } else {
        int rc = halo_exchange(c, c->d_p.p);
        if (rc) return rc;
    }
    launch_spmv<true>(c, c->d_p.p, c->d_q.p + o6, c->d_p.p + o6, red, fin, pw);
    if (fin == FIN_RED) {
        FS_NCCL(c, nccl().AllReduce(red, red, 1, ncclDouble, ncclSum, (ncclComm_t)c->comm, c->stream));
        k_finalize<<<1, 1, 0, c->stream>>>(c->d_state.p, red, 1);
    }
    if (PC == 3) {  // z needs the lattice restriction of the updated r: update, restrict, coarse levels, prolong
        k_update_xr<256><<<vg, 256, 0, c->stream>>>(c->n_own, c->d_x.p + o6, c->d_r.p + o6, c->d_p.p + o6, c->d_q.p + o6,
                                                    c->d_partials.p, c->d_counter.p, c->d_state.p, red + 5);
        mark(1);
        int rc = ml_enqueue_apply(c, false, red + 4, fin, vg);
        if (rc) return rc;
        mark(5);
        if (fin == FIN_RED) {
            FS_NCCL(c, nccl().AllReduce(red + 4, red + 4, 2, ncclDouble, ncclSum, (ncclComm_t)c->comm, c->stream));
            k_finalize<<<1, 1, 0, c->stream>>>(c->d_state.p, red + 4, 2);
        }
        k_direction<256><<<vg, 256, 0, c->stream>>>(c->n_own, c->d_z.p + o6, c->d_p.p + o6, c->d_state.p, pw, c->d_counter.p, PushArgs{});
        mark(6);
        return FS_OK;
    }
    k_update<PC == 3 ? 1 : PC, NORM, 256><<<vg, 256, 0, c->stream>>>(c->n_own, c->d_x.p + o6, c->d_r.p + o6, c->d_p.p + o6,
                                                       c->d_q.p + o6, c->d_z.p + o6, c->d_minv.p, c->d_partials.p,
                                                       c->d_counter.p, c->d_state.p, red + 4, fin, pw);
    if (fin == FIN_RED) {
        FS_NCCL(c, nccl().AllReduce(red + 4, red + 4, 2, ncclDouble, ncclSum, (ncclComm_t)c->comm, c->stream));
        k_finalize<<<1, 1, 0, c->stream>>>(c->d_state.p, red + 4, 2);
    }
    k_direction<256><<<vg + push.n_push_blocks, 256, 0, c->stream>>>(c->n_own, c->d_z.p + o6, c->d_p.p + o6, c->d_state.p, pw, c->d_counter.p, push);
    return FS_OK;
}

template <int PC, int NORM>
static int run_pcg(fs_context *c, const fs_solve_opts *o, fs_solve_info *info)
{
    cudaStream_t st = c->stream;
    const bool single = (c->world == 1), peer = !single && c->peer_ready && PC != 3;
    const int fin = single ? FIN_INLINE : (peer ? FIN_PEER : FIN_RED);
    PeerWin *pw = peer ? c->d_pw.p : nullptr;
    const int64_t o6 = 6 * c->own_lo;
    const int sg = spmv_grid(c), vg = vec_grid(c);
    const int maxgrid = std::max(std::max(sg, vg), c->sm_count * c->sell_blocks_per_sm);
    if (c->d_partials.n < (size_t)maxgrid * 4 + 16) FS_CUDA(c, c->d_partials.alloc((size_t)maxgrid * 4 + 16));
    double *red = c->d_partials.p + (size_t)maxgrid * 4;  // 16 spare doubles behind the partials

    CgState h = {};
    h.tol2 = o->rtol * o->rtol;
    h.max_its = o->max_its;
    h.status = FS_ERR_NOT_CONVERGED;
    FS_CUDA(c, cudaMemcpyAsync(c->d_state.p, &h, sizeof h, cudaMemcpyHostToDevice, st));
    FS_CUDA(c, cudaMemsetAsync(c->d_counter.p, 0, 4 * sizeof(unsigned int), st));
    if (!o->warm_start || !c->have_solution)
        FS_CUDA(c, cudaMemsetAsync(c->d_x.p, 0, sizeof(double) * 6 * c->n_local, st));

    FS_CUDA(c, cudaEventRecord(c->ev0, st));
    // r = b - A x0
    int rc = spmv_once(c, c->d_x.p, c->d_q.p);
    if (rc) return rc;
    if (PC == 3) {
        k_init_r<256><<<vg, 256, 0, st>>>(c->n_own, c->d_b.p + o6, c->d_q.p + o6, c->d_r.p + o6, c->d_partials.p, c->d_counter.p, red + 9);
        rc = ml_enqueue_apply(c, true, red + 8, fin, vg);
        if (rc) return rc;
    } else
        k_init<PC == 3 ? 1 : PC, NORM, 256><<<vg, 256, 0, st>>>(c->n_own, c->d_b.p + o6, c->d_q.p + o6, c->d_r.p + o6, c->d_z.p + o6,
                                                                 c->d_p.p + o6, c->d_minv.p, c->d_partials.p, c->d_counter.p,
                                                                 c->d_state.p, red + 8, fin, pw);
    if (fin == FIN_RED) {
        FS_NCCL(c, nccl().AllReduce(red + 8, red + 8, 3, ncclDouble, ncclSum, (ncclComm_t)c->comm, st));
        k_finalize<<<1, 1, 0, st>>>(c->d_state.p, red + 8, 0);
    } else if (fin == FIN_PEER) {
        k_finalize_init_peer<<<1, 32, 0, st>>>(c->d_state.p, pw);
        // halo of the first direction p = z; every later one is pushed by the k_direction that forms it
        if (c->push_foldable)
            k_halo_push<<<std::max(1u, nblk(3 * c->send_total, 256)), 256, 0, st>>>(
                pw, c->send_total, c->d_send_idx.p, c->d_push_peer.p, c->d_push_dst.p, c->d_p.p, c->d_state.p, c->d_counter.p + 1);
    }
    // The iteration is captured once into a CUDA graph of GRAPH_ITERS iterations and replayed; kernels
    // past convergence (or past max_its) see done != 0 and return, so replaying whole graphs is exact.
    if (PC == 3 && getenv("FS_ML_PROFILE")) {   // lab: eager iterations, stage times from events (tools/ml_stage_profile.py)
        fs_context::StageProf &pf = c->prof;
        if (!pf.ev[0])
            for (cudaEvent_t &e : pf.ev) FS_CUDA(c, cudaEventCreate(&e));
        pf.on = true;
        for (;;) {
            FS_CUDA(c, cudaMemcpyAsync(c->h_state, c->d_state.p, sizeof(CgState), cudaMemcpyDeviceToHost, st));
            FS_CUDA(c, cudaStreamSynchronize(st));
            if (c->h_state->done) break;
            rc = enqueue_iteration<PC, NORM>(c, red, sg, vg);
            if (rc) { pf.on = false; return rc; }
            FS_CUDA(c, cudaStreamSynchronize(st));
            for (int k = 0; k < 6; k++) {
                float ms = 0.f;
                FS_CUDA(c, cudaEventElapsedTime(&ms, pf.ev[k], pf.ev[k + 1]));
                pf.ms[k] += ms;
            }
            for (int k = 0; k < 2; k++) {   // the two visits of the second lattice level from the first
                float ms = 0.f;
                if (cudaEventElapsedTime(&ms, pf.ev[7 + 2 * k], pf.ev[8 + 2 * k]) == cudaSuccess) pf.ms[6 + k] += ms;
                else cudaGetLastError();
            }
            pf.n++;
        }
        pf.on = false;
    } else {
        constexpr int GRAPH_ITERS = (PC == 3) ? 1 : 8;  // a multilevel iteration is ~100 launches already
        const int key = PC * 2 + NORM + (c->sell_active ? 8 * (1 + c->sell_kind) : 0) + (peer ? 64 : 0);
        if (!c->cg_graph_exec || c->cg_graph_key != key || c->cg_graph_red != red) {
            PhaseTimer tmc("run_pcg");
            if (c->cg_graph_exec) cudaGraphExecDestroy(c->cg_graph_exec);
            c->cg_graph_exec = nullptr;
            cudaGraph_t graph = nullptr;
            FS_CUDA(c, cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            rc = FS_OK;
            for (int k = 0; k < GRAPH_ITERS && rc == FS_OK; k++) rc = enqueue_iteration<PC, NORM>(c, red, sg, vg);
            cudaError_t ce = cudaStreamEndCapture(st, &graph);
            if (rc) {
                if (graph) cudaGraphDestroy(graph);
                return rc;
            }
            FS_CUDA(c, ce);
            FS_CUDA(c, cudaGraphInstantiate(&c->cg_graph_exec, graph, 0));
            FS_CUDA(c, cudaGraphDestroy(graph));
            tmc.lap("graph capture + instantiate");
            c->cg_graph_key = key;
            c->cg_graph_red = red;
        }
        const int batch = o->check_every > 0 ? o->check_every : (PC == 3 ? 4 : 64);
        for (;;) {
            FS_CUDA(c, cudaMemcpyAsync(c->h_state, c->d_state.p, sizeof(CgState), cudaMemcpyDeviceToHost, st));
            FS_CUDA(c, cudaStreamSynchronize(st));
            if (c->h_state->done) break;
            int64_t left = c->h_state->max_its - c->h_state->iter;
            int64_t n = std::min<int64_t>(batch, std::max<int64_t>(left, 1));
            for (int64_t k = 0; k < n; k += GRAPH_ITERS) FS_CUDA(c, cudaGraphLaunch(c->cg_graph_exec, st));
            FS_CUDA(c, cudaGetLastError());
        }
    }
    FS_CUDA(c, cudaEventRecord(c->ev1, st));
    FS_CUDA(c, cudaStreamSynchronize(st));
    const CgState &s = *c->h_state;
    if (s.nrm2 < 0.0) FS_CUDA(c, cudaMemsetAsync(c->d_x.p, 0, sizeof(double) * 6 * c->n_local, st));
    // a later warm start (the default, fsp.cpp:271) must not begin from the iterate of a broken-down or timed-out solve
    c->have_solution = (s.status == FS_OK || s.status == FS_ERR_NOT_CONVERGED) && std::isfinite(s.nrm2) && std::isfinite(s.rz);
    if (info) {
        info->iterations = s.iter;
        info->rel_residual = s.nrm2 < 0.0 ? 0.0 : sqrt(s.nrm2 / s.bnorm2);
        info->status = s.status;
        info->spmv_ms = 0.f;
        FS_CUDA(c, cudaEventElapsedTime(&info->solve_ms, c->ev0, c->ev1));
    }
    if (s.status == FS_ERR_COMM) return fail(c, FS_ERR_COMM, "a wait for another GPU's halo or partial sums timed out (peer windows)");
    if (s.status == FS_ERR_BREAKDOWN) return fail(c, FS_ERR_BREAKDOWN, "CG breakdown: p.Ap <= 0 (matrix not positive definite)");
    if (s.status == FS_ERR_NOT_CONVERGED) return fail(c, FS_ERR_NOT_CONVERGED, "CG did not reach rtol within max_its");
    return FS_OK;
}

// z = M^-1 r once, on device vectors in the local layout (tests of the multilevel preconditioner): r in d_r, z in d_z
int pc_apply_mlrbm_once(fs_context *c)
{
    int rc = solver_prepare(c, FS_PC_MLRBM);
    if (rc) return rc;
    rc = ml_enqueue_apply(c, true, nullptr, FIN_RED, vec_grid(c));
    if (rc) return rc;
    FS_CUDA(c, cudaGetLastError());
    return FS_OK;
}

// the SpMV kernel alone (no halo exchange, no reduction), reps launches between two events
int spmv_kernel_time(fs_context *c, int reps, float *ms_per_launch, float *ms_on_p)
{
    int rc = spmv_format_prepare(c);
    if (rc) return rc;
    rc = halo_exchange(c, c->d_b.p);
    if (rc) return rc;
    const int64_t o6 = 6 * c->own_lo;
    launch_spmv<false>(c, c->d_b.p, c->d_q.p + o6, nullptr, nullptr, FIN_RED, nullptr, false);
    FS_CUDA(c, cudaEventRecord(c->ev0, c->stream));
    for (int i = 0; i < reps; i++) launch_spmv<false>(c, c->d_b.p, c->d_q.p + o6, nullptr, nullptr, FIN_RED, nullptr, false);
    FS_CUDA(c, cudaEventRecord(c->ev1, c->stream));
    FS_CUDA(c, cudaStreamSynchronize(c->stream));
    FS_CUDA(c, cudaGetLastError());
    float ms = 0.f;
    FS_CUDA(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    *ms_per_launch = ms / reps;
    if (ms_on_p) {  // same kernel reading the direction vector p (inside the cudaIpc window on the peer path)
        FS_CUDA(c, cudaEventRecord(c->ev0, c->stream));
        for (int i = 0; i < reps; i++) launch_spmv<false>(c, c->d_p.p, c->d_q.p + o6, nullptr, nullptr, FIN_RED, nullptr, false);
        FS_CUDA(c, cudaEventRecord(c->ev1, c->stream));
        FS_CUDA(c, cudaStreamSynchronize(c->stream));
        FS_CUDA(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        *ms_on_p = ms / reps;
    }
    return FS_OK;
}

int solver_run(fs_context *c, const fs_solve_opts *o, fs_solve_info *info)
{
    if (!c->rhs_ready) return fail(c, FS_ERR_STATE, "no right-hand side: set loads first");
    PhaseTimer tm("solver_run");
    int rc = solver_prepare(c, o->pc);
    if (rc) return rc;
    if (tm.on) cudaStreamSynchronize(c->stream);
    tm.lap("solver_prepare (format, preconditioner)");
    const int nt = o->norm_type ? 1 : 0;
    switch (o->pc * 2 + nt) {
    case 0: return run_pcg<0, 0>(c, o, info);
    case 1: return run_pcg<0, 1>(c, o, info);
    case 2: return run_pcg<1, 0>(c, o, info);
    case 3: return run_pcg<1, 1>(c, o, info);
    case 4: return run_pcg<2, 0>(c, o, info);
    case 5: return run_pcg<2, 1>(c, o, info);
    case 6: return run_pcg<3, 0>(c, o, info);
    default: return fail(c, FS_ERR_ARG, "FS_PC_MLRBM tests convergence in the unpreconditioned norm only");
    }
}

}  // namespace fs

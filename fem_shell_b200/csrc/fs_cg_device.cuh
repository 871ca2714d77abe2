// fs_cg_device.cuh -- device pieces shared by the CG kernels (fs_solver.cu) and the SpMV variants
// (fs_sell.cuh): the deterministic grid reduction and the scalar recurrences of the iteration.
#pragma once
#include "fs_context.hpp"
#include "fs_peer.cuh"

namespace fs {

// ---------------------------------------------------------------------------------------------
// deterministic grid reduction of NV values; returns true in the block that arrives last, with
// the totals in out[] (valid for thread 0 of that block)
// ---------------------------------------------------------------------------------------------
template <int NV, int BLOCK>
__device__ __forceinline__ bool grid_reduce(double (&v)[NV], double *partials, unsigned int *counter,
                                            double (&out)[NV])
{
    __shared__ double s_red[NV][BLOCK / 32];
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; k++) {
        double x = v[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) s_red[k][warp] = x;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NV; k++) {
            double x = 0.0;
            for (int w = 0; w < BLOCK / 32; w++) x += s_red[k][w];
            partials[(size_t)blockIdx.x * NV + k] = x;
        }
        __threadfence();
        unsigned int ticket = atomicInc(counter, gridDim.x - 1);  // wraps back to 0 for the next use
        s_last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return false;
    __threadfence();
    // fixed-order sum of the per-block partials
#pragma unroll
    for (int k = 0; k < NV; k++) {
        double x = 0.0;
        for (unsigned int b = threadIdx.x; b < gridDim.x; b += BLOCK) x += partials[(size_t)b * NV + k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        __syncthreads();
        if (lane == 0) s_red[k][warp] = x;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NV; k++) {
            double x = 0.0;
            for (int w = 0; w < BLOCK / 32; w++) x += s_red[k][w];
            out[k] = x;
        }
    }
    return true;
}

// ---- scalar recurrences (run by one thread) --------------------------------------------------
__device__ __forceinline__ void finalize_pq(CgState *s, double pq)
{
    s->pq = pq;
    s->alpha = s->rz / pq;
    if (!(pq > 0.0)) {  // not SPD (SURVEY.md section 7 "SPD is empirical")
        s->status = FS_ERR_BREAKDOWN;
        s->done = 1;
    }
}

__device__ __forceinline__ void finalize_update(CgState *s, double rz_new, double nrm2)
{
    s->beta = rz_new / s->rz;
    s->rz = rz_new;
    s->nrm2 = nrm2;
    s->iter += 1;
    if (nrm2 <= s->tol2 * s->bnorm2) {
        s->status = FS_OK;
        s->done = 1;
    } else if (s->iter >= s->max_its) {
        s->status = FS_ERR_NOT_CONVERGED;
        s->done = 1;
    }
}

__device__ __forceinline__ void finalize_init(CgState *s, double rz, double nrm2, double bnorm2)
{
    s->rz = rz;
    s->nrm2 = nrm2;
    s->bnorm2 = bnorm2;
    s->iter = 0;
    s->status = FS_ERR_NOT_CONVERGED;
    s->done = 0;
    if (bnorm2 == 0.0) {  // b = 0 -> x = 0 (host zeroes x when it sees nrm2 < 0)
        s->bnorm2 = 1.0;
        s->nrm2 = -1.0;
        s->status = FS_OK;
        s->done = 1;
    } else if (nrm2 <= s->tol2 * bnorm2) {
        s->status = FS_OK;
        s->done = 1;
    } else if (s->max_its <= 0) {
        s->done = 1;
    }
}

// what the block that finished a grid reduction does with the totals.  fin_mode: 0 leave them in red[] (an
// ncclAllReduce + k_finalize follow), 1 single rank: advance the recurrence here, 2 push them to every
// rank's mailbox (fs_peer.cuh); the consuming kernel completes the sum
enum { FIN_RED = 0, FIN_INLINE = 1, FIN_PEER = 2 };
template <int NV>
__device__ __forceinline__ void finish_dot(const double (&out)[NV], double *red, int fin_mode, CgState *state, PeerWin *pw)
{
#pragma unroll
    for (int k = 0; k < NV; k++) red[k] = out[k];
    if (fin_mode == FIN_PEER) peer_red_push<NV>(pw, out);
    else if (fin_mode == FIN_INLINE) {
        if (NV == 1) finalize_pq(state, out[0]);
        else if (NV == 2) finalize_update(state, out[0], out[NV > 1 ? 1 : 0]);
        else finalize_init(state, out[0], out[NV > 1 ? 1 : 0], out[NV > 2 ? 2 : 0]);
    }
}

__device__ __forceinline__ void load6(const double *p, double v[6])
{
    const double2 *q = reinterpret_cast<const double2 *>(p);
#pragma unroll
    for (int h = 0; h < 3; h++) {
        double2 t = q[h];
        v[2 * h] = t.x;
        v[2 * h + 1] = t.y;
    }
}
__device__ __forceinline__ void store6(double *p, const double v[6])
{
    double2 *q = reinterpret_cast<double2 *>(p);
#pragma unroll
    for (int h = 0; h < 3; h++) q[h] = make_double2(v[2 * h], v[2 * h + 1]);
}

}  // namespace fs

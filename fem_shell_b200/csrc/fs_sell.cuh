// fs_sell.cuh -- zero-compacted SpMV format for matrices whose 6x6 node blocks share a sparse pattern.
//
// libMesh pre-allocates (and the reference's add_matrix fills, fs.cpp:1230) a dense 6x6 block for every
// node pair sharing an element, explicit zeros included.  For a flat shell lying in a coordinate plane
// (every meshGen plate, src/meshgen/main_all.cpp:141-160) the local->global rotation (fs.cpp:1061-1110)
// is a signed permutation, so membrane, bending and drilling DOFs never couple and 22 of the 36 entries
// of EVERY block are exact zeros.  The parity format (d_vals, what fs_export_csr returns) keeps them; the
// SpMV does not have to stream them.  After each values pass the 36-bit union pattern of all blocks is
// measured on the device; when it fits one of the masks below, the non-zero positions are copied into a
// sliced-ELL layout and the CG iteration runs on that (same products, same column order, exact zeros
// skipped: results are those of a sequential CSR row sum).
//
// Layout ("SELL-32 by node"): a slice = 32 consecutive owned block rows, one per lane; dmax = largest
// number of blocks of any row in the slice; sptr[s] = sum of dmax over earlier slices.
//   adj [32*(sptr[s]+slot) + lane]                  local column node of the lane's slot-th block (padding: own node)
//   vals[32*(NZ*(sptr[s]+slot) + item) + lane]      item = rank of (a,b) among the set bits of the mask (row-major)
// so every load of a warp is one contiguous 256-byte (values) or 128-byte (columns) line, a lane owns a
// whole block row, and no cross-lane reduction is needed.
#pragma once
#include "fs_cg_device.cuh"
#include "fs_peer.cuh"

namespace fs {

// launch shape of k_spmv_sell (tools/sell_lab.cu)
constexpr int SELL_BLOCK = 128;
constexpr int SELL_MINB = 4;

__host__ __device__ constexpr unsigned long long sell_bit(int a, int b) { return 1ull << (6 * a + b); }
__host__ __device__ constexpr unsigned long long sell_group(int i, int j, int k)  // rows/cols {i,j,k} fully coupled (k < 0: pair, j < 0: single)
{
    unsigned long long m = 0;
    const int g[3] = {i, j, k};
    for (int p = 0; p < 3; p++)
        for (int q = 0; q < 3; q++)
            if (g[p] >= 0 && g[q] >= 0) m |= sell_bit(g[p], g[q]);
    return m;
}
// shell in the xy plane: membrane (u,v), bending (w,tx,ty), drilling tz -- and the two other coordinate planes
constexpr unsigned long long SELL_MASK_XY = sell_group(0, 1, -1) | sell_group(2, 3, 4) | sell_group(5, -1, -1);
constexpr unsigned long long SELL_MASK_XZ = sell_group(0, 2, -1) | sell_group(1, 3, 5) | sell_group(4, -1, -1);
constexpr unsigned long long SELL_MASK_YZ = sell_group(1, 2, -1) | sell_group(0, 4, 5) | sell_group(3, -1, -1);
constexpr unsigned long long SELL_MASK_FULL = (1ull << 36) - 1;

__host__ __device__ constexpr int sell_popcount(unsigned long long m)
{
    int n = 0;
    for (; m; m &= m - 1) n++;
    return n;
}
__host__ __device__ constexpr int sell_item(unsigned long long mask, int a, int b) { return sell_popcount(mask & (sell_bit(a, b) - 1)); }
__host__ __device__ constexpr bool sell_uses_col(unsigned long long mask, int b)
{
    for (int a = 0; a < 6; a++)
        if (mask & sell_bit(a, b)) return true;
    return false;
}

// ---------------------------------------------------------------------------------------------
// q = A p on the compacted matrix (+ partial p.q).  Lane = block row; slots of a slice are walked in
// column order, two at a time so that 2*NZ value loads are in flight per lane.
// ---------------------------------------------------------------------------------------------
// PEER = the halo segments of x are written by the neighbouring GPUs while this kernel is already resident (it
// spins on their stamps first, fs_peer.cuh).  x is then NOT read-only for the kernel's lifetime: no __restrict__ /
// ld.global.nc on it, and halo blocks are loaded with ld.global.cg (L2 is the point of coherence for NVLink
// writes; an L1 line could predate them).  Owned blocks of x are only written by earlier kernels of this stream.
template <bool PEER>
__device__ __forceinline__ double2 ld_x2(const double2 *p, bool halo)
{
    if (PEER) return halo ? __ldcg(p) : __ldg(p);   // owned blocks are read-only for the kernel's lifetime: ld.global.nc is legal
    return __ldg(p);
}

// plane frame of a planar shell in general position (fs_context.hpp): rows x^, y^, n
struct PlaneQ {
    double m[3][3];
};

// one slice: the lane's block row times x; returns the lane's contribution to p.q.  ROT: the stored blocks are
// Q~ K Q~^T -- every x block is rotated into the plane frame (Q per translation / rotation triple) and the finished
// row back out (Q^T): 18 + 18 FMAs on a kernel that waits for HBM.
template <unsigned long long MASK, bool WITH_DOT, bool PEER, int UNROLL, bool ROT>
__device__ __forceinline__ double sell_slice(int s, int lane, int n_own, int own_lo, const int32_t *__restrict__ sptr,
                                             const int32_t *__restrict__ adj, const double *__restrict__ vals, const double *x,
                                             double *__restrict__ y_own, const double *x_own, const PlaneQ &q)
{
    constexpr int NZ = sell_popcount(MASK);
    const int s0 = sptr[s], dmax = sptr[s + 1] - s0;
    const int32_t *aj = adj + 32 * (size_t)s0 + lane;
    const double *v = vals + 32 * (size_t)NZ * s0 + lane;
    double acc[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll UNROLL
    for (int slot = 0; slot < dmax; slot++) {
        const int col = aj[32 * slot];
        const bool halo = PEER && (unsigned)(col - own_lo) >= (unsigned)n_own;
        const double2 *xp = reinterpret_cast<const double2 *>(x + 6 * (size_t)col);
        double xv[6];
#pragma unroll
        for (int h = 0; h < 3; h++)
            if (ROT || sell_uses_col(MASK, 2 * h) || sell_uses_col(MASK, 2 * h + 1)) {
                const double2 t = ld_x2<PEER>(xp + h, halo);
                xv[2 * h] = t.x;
                xv[2 * h + 1] = t.y;
            }
        if (ROT) {
            double r[6];
#pragma unroll
            for (int i = 0; i < 3; i++) {
                r[i] = q.m[i][0] * xv[0] + q.m[i][1] * xv[1] + q.m[i][2] * xv[2];
                r[3 + i] = q.m[i][0] * xv[3] + q.m[i][1] * xv[4] + q.m[i][2] * xv[5];
            }
#pragma unroll
            for (int i = 0; i < 6; i++) xv[i] = r[i];
        }
        const double *vs = v + 32 * (size_t)NZ * slot;
#pragma unroll
        for (int a = 0; a < 6; a++)
#pragma unroll
            for (int b = 0; b < 6; b++)
                if (MASK & sell_bit(a, b)) acc[a] += __ldcs(vs + 32 * sell_item(MASK, a, b)) * xv[b];
    }
    if (ROT) {
        double r[6];
#pragma unroll
        for (int i = 0; i < 3; i++) {
            r[i] = q.m[0][i] * acc[0] + q.m[1][i] * acc[1] + q.m[2][i] * acc[2];
            r[3 + i] = q.m[0][i] * acc[3] + q.m[1][i] * acc[4] + q.m[2][i] * acc[5];
        }
#pragma unroll
        for (int i = 0; i < 6; i++) acc[i] = r[i];
    }
    const int p = 32 * s + lane;
    double dot = 0.0;
    if (p < n_own) {
        store6(y_own + 6 * (size_t)p, acc);
        if (WITH_DOT) {
            double pv[6];
            load6(x_own + 6 * (size_t)p, pv);
#pragma unroll
            for (int a = 0; a < 6; a++) dot += acc[a] * pv[a];
        }
    }
    return dot;
}

// PEER: hflag[s] != 0 marks the slices whose rows read halo blocks (a few per cent of a strip: its first and last
// node rows).  A warp waits for the neighbours' stamps only when it reaches its first such slice (the push left with
// the leading blocks of the previous k_direction, so the wait is usually over before it starts) and loads that
// slice's halo blocks past L1; every other slice runs the single-rank loop -- owned blocks of x are read-only for
// this kernel's lifetime.  Slices are walked in storage order: measured on 2 GPUs (profiles/r02h), moving the halo
// slices behind a mid-kernel wait cost one extra slice time (~9 us of a 176 us kernel) in the tail.
template <unsigned long long MASK, bool WITH_DOT, int BLOCK, int MINB, bool PEER = false, bool ROT = false, int UNROLL = 2>
__global__ void __launch_bounds__(BLOCK, MINB)
k_spmv_sell(int n_own, int own_lo, int n_slices, const int32_t *__restrict__ sptr, const int32_t *__restrict__ adj,
            const double *__restrict__ vals, const double *x, double *__restrict__ y_own,
            const double *x_own, double *partials, unsigned int *counter, CgState *state,
            double *red, int fin_mode, PeerWin *pw, const int32_t *__restrict__ hflag, const __grid_constant__ PlaneQ q)
{
    if (WITH_DOT ? state->done : (state && state->done)) return;  // without the dot product: checked only when a state is passed
    const int lane = threadIdx.x & 31;
    const int gw = blockIdx.x * (BLOCK / 32) + (threadIdx.x >> 5);
    const int nw = gridDim.x * (BLOCK / 32);
    double dot = 0.0;
    if (PEER) {
        peer_kern_begin(pw, 0);
        bool waited = false;
        int f_next = gw < n_slices ? hflag[gw] : 0;
        for (int s = gw; s < n_slices; s += nw) {
            const int f = f_next;
            if (s + nw < n_slices) f_next = hflag[s + nw];
            if (f) {   // warp-uniform
                if (!waited) {
                    if (!peer_halo_wait_warp(pw, lane)) peer_fail(state);   // the solve ends as FS_ERR_COMM; this product is discarded
                    waited = true;
                }
                dot += sell_slice<MASK, WITH_DOT, true, UNROLL, ROT>(s, lane, n_own, own_lo, sptr, adj, vals, x, y_own, x_own, q);
            } else
                dot += sell_slice<MASK, WITH_DOT, false, UNROLL, ROT>(s, lane, n_own, own_lo, sptr, adj, vals, x, y_own, x_own, q);
        }
    } else {
        for (int s = gw; s < n_slices; s += nw) dot += sell_slice<MASK, WITH_DOT, false, UNROLL, ROT>(s, lane, n_own, own_lo, sptr, adj, vals, x, y_own, x_own, q);
    }
    if (WITH_DOT) {
        double vv[1] = {dot}, out[1];
        if (grid_reduce<1, BLOCK>(vv, partials, counter, out) && threadIdx.x == 0) {
            finish_dot<1>(out, red, fin_mode, state, pw);
            if (PEER) peer_kern_end(pw, 0);
        }
    }
}

// per slice: 1 when a row of the slice reads a block outside the owned range (a halo block)
static __global__ void k_sell_halo_flags(int n_own, int own_lo, int n_slices, const int32_t *__restrict__ sptr,
                                         const int32_t *__restrict__ adj, int32_t *__restrict__ flag)
{
    const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (s >= n_slices) return;
    const int s0 = sptr[s], dmax = sptr[s + 1] - s0;
    bool any = false;
    for (int slot = 0; slot < dmax; slot++) any |= (unsigned)(adj[32 * (size_t)(s0 + slot) + lane] - own_lo) >= (unsigned)n_own;
    any = __any_sync(0xffffffffu, any);
    if (lane == 0) flag[s] = any ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------
// format build
// ---------------------------------------------------------------------------------------------
// union pattern of all 6x6 blocks: bit 6a+b set when any block has a non-zero (a,b).  Same access
// pattern as the full SpMV (warp = block row, lane = one double2 of each scalar row).
static __global__ void __launch_bounds__(256) k_sell_detect(int n_own, const int32_t *__restrict__ nptr,
                                                     const double *__restrict__ vals, unsigned long long *mask_out)
{
    const int lane = threadIdx.x & 31;
    const int gw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int nw = gridDim.x * (blockDim.x >> 5);
    unsigned long long m = 0;
    for (int p = gw; p < n_own; p += nw) {
        const int b0 = nptr[p], L2 = 3 * (nptr[p + 1] - b0);
        const double2 *base = reinterpret_cast<const double2 *>(vals + (size_t)36 * b0);
        for (int l = lane; l < L2; l += 32) {
            const int h = l % 3;
#pragma unroll
            for (int a = 0; a < 6; a++) {
                const double2 t = __ldcs(base + (size_t)a * L2 + l);
                if (t.x != 0.0) m |= sell_bit(a, 2 * h);
                if (t.y != 0.0) m |= sell_bit(a, 2 * h + 1);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m |= __shfl_xor_sync(0xffffffffu, m, o);
    if (lane == 0 && m) atomicOr(mask_out, m);
}

static __global__ void k_sell_dmax(int n_own, int n_slices, const int32_t *__restrict__ nptr, int32_t *dmax)
{
    const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (s >= n_slices) return;
    const int p = 32 * s + lane;
    int d = p < n_own ? nptr[p + 1] - nptr[p] : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d = max(d, __shfl_xor_sync(0xffffffffu, d, o));
    if (lane == 0) dmax[s] = d;
    if (s == 0 && lane == 0) dmax[n_slices] = 0;
}

// warp = slice, lane = block row: copies the masked entries of the parity format into the sliced layout
static __global__ void __launch_bounds__(256) k_sell_fill(int n_own, int n_slices, int own_lo, unsigned long long mask,
                                                   const int32_t *__restrict__ nptr, const int32_t *__restrict__ nadj,
                                                   const double *__restrict__ full, const int32_t *__restrict__ sptr,
                                                   int32_t *__restrict__ adj, double *__restrict__ vals, int nz,
                                                   int write_adj)
{
    const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (s >= n_slices) return;
    const int p = 32 * s + lane;
    const bool live = p < n_own;
    const int b0 = live ? nptr[p] : 0, deg = live ? nptr[p + 1] - b0 : 0;
    const int s0 = sptr[s], dmax = sptr[s + 1] - s0;
    const int self = own_lo + (live ? p : 0);
    const double *row = full + (size_t)36 * b0;
    for (int slot = 0; slot < dmax; slot++) {
        const bool have = slot < deg;
        if (write_adj) adj[32 * (size_t)(s0 + slot) + lane] = have ? nadj[b0 + slot] : self;
        double *dst = vals + 32 * ((size_t)nz * (s0 + slot)) + lane;
        int item = 0;
        for (int a = 0; a < 6; a++)
            for (int b = 0; b < 6; b++)
                if (mask & sell_bit(a, b)) {
                    dst[32 * (size_t)item] = have ? row[(size_t)a * 6 * deg + 6 * slot + b] : 0.0;
                    item++;
                }
    }
}

// Same copy as k_sell_fill, but every byte moves coalesced: block = one slice (32 block rows, contiguous in the
// parity array), the four warps stream the rows into shared memory (row stride odd -> the transposed reads below
// are bank-conflict free), then lane = row writes whole 256-byte lines of the sliced-ELL layout.  The union
// pattern of the blocks is measured on the way (what k_sell_detect does in a pass of its own), so a re-assembly
// costs ONE read of the parity values instead of two scattered ones (profiles/r01h: detect 0.42 ms + fill 1.46 ms
// per values pass on c2 before this kernel).
constexpr int SELL_FILL_THREADS = 256;

static __global__ void __launch_bounds__(SELL_FILL_THREADS) k_sell_fill_t(int n_own, int own_lo, unsigned long long mask,
                                                                 const int32_t *__restrict__ nptr, const int32_t *__restrict__ nadj,
                                                                 const double *__restrict__ full, const int32_t *__restrict__ sptr,
                                                                 int32_t *__restrict__ adj, double *__restrict__ vals, int nz,
                                                                 int write_adj, unsigned long long *mask_out)
{
    constexpr int NW = SELL_FILL_THREADS / 32;
    extern __shared__ double sm_rows[];
    __shared__ int s_ab[36];
    const int s = blockIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int s0 = sptr[s], dmax = sptr[s + 1] - s0;
    const int S = (36 * dmax) | 1;
    const int p = 32 * s + lane;
    const bool live = p < n_own;
    const int b0 = live ? nptr[p] : 0, deg = live ? nptr[p + 1] - b0 : 0, L = 6 * deg;
    if (threadIdx.x < 36) {  // item -> (a, b) of the mask's set bits, row-major
        const int a = threadIdx.x / 6, b = threadIdx.x % 6;
        if (mask & sell_bit(a, b)) s_ab[sell_popcount(mask & (sell_bit(a, b) - 1))] = (a << 4) | b;
    }
    unsigned long long m = 0;
#pragma unroll
    for (int rr = 0; rr < 32 / NW; rr++) {  // the rows of a warp are independent: all their loads are in flight together
        const int r = NW * rr + w;
        const int rb0 = __shfl_sync(0xffffffffu, b0, r), rL = __shfl_sync(0xffffffffu, L, r);
        const double2 *src2 = reinterpret_cast<const double2 *>(full + (size_t)36 * rb0);  // 288-byte multiples: aligned
        double *dst = sm_rows + (size_t)r * S;
        const int n2 = 3 * rL;  // double2 items of the row (36 * deg / 2)
        for (int base = 0; base < n2; base += 6 * 32) {
            double2 buf[6];
#pragma unroll
            for (int i = 0; i < 6; i++) {
                const int idx = base + 32 * i + lane;
                buf[i] = idx < n2 ? __ldcs(src2 + idx) : make_double2(0.0, 0.0);
            }
#pragma unroll
            for (int i = 0; i < 6; i++) {
                const int idx = base + 32 * i + lane;
                if (idx < n2) {
                    const int k = 2 * idx;  // k even, rL even: k and k+1 lie in the same scalar row a, b = k % 6 does not wrap
                    dst[k] = buf[i].x;
                    dst[k + 1] = buf[i].y;
                    const int a = (k >= rL) + (k >= 2 * rL) + (k >= 3 * rL) + (k >= 4 * rL) + (k >= 5 * rL), b = k % 6;
                    if (buf[i].x != 0.0) m |= sell_bit(a, b);
                    if (buf[i].y != 0.0) m |= sell_bit(a, b + 1);
                }
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m |= __shfl_xor_sync(0xffffffffu, m, o);
    if (lane == 0 && (m & ~mask)) atomicOr(mask_out, m);  // only a pattern outside the mask needs reporting
    __syncthreads();
    const int self = own_lo + (live ? p : 0);
    if (write_adj)
        for (int slot = w; slot < dmax; slot += NW) adj[32 * (size_t)(s0 + slot) + lane] = slot < deg ? nadj[b0 + slot] : self;
    const double *row = sm_rows + (size_t)lane * S;
    double *base = vals + 32 * ((size_t)nz * s0) + lane;
    const int n_lines = dmax * nz;  // one 256-byte line per (slot, item)
#pragma unroll 4
    for (int q = w; q < n_lines; q += NW) {
        const int slot = q / nz, ab = s_ab[q - slot * nz];
        const double v = slot < deg ? row[(ab >> 4) * L + 6 * slot + (ab & 15)] : 0.0;
        __stcs(base + 32 * (size_t)q, v);
    }
}

}  // namespace fs

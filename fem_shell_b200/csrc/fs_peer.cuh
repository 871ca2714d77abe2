// fs_peer.cuh -- the CG iteration's two exchanges done by the kernels themselves over NVLink peer memory.
//
// The reference's Krylov loop (KSPSolve, fs.cpp:138) costs PETSc one VecScatter (halo of p) and two
// MPI_Allreduce per iteration.  With one process per GPU on an NVSwitch box every rank maps a small
// "window" of every other rank (cudaIpc): [mailbox | direction vector p].  Then
//   * the halo of p is PUSHED by its owner straight into the neighbours' halo segments of p (k_halo_push,
//     remote 128-bit stores) and stamped; the neighbour's SpMV waits for the stamp at its first instruction;
//   * a dot product is finished by the producing kernel: its last block stores the rank's partial sums into
//     every rank's mailbox; the consuming kernel (k_update for p.Ap, k_direction for r.z and the norm) waits
//     for all of them and adds the partials in rank order -- every rank gets bit-identical alpha/beta,
//     deterministic, no reduction kernel, no NCCL launch in the loop.  The partials travel as self-certifying
//     8-byte words {32 bits of the double | 32-bit stamp} (the "LL" trick of NCCL's low-latency protocol): an
//     8-byte store is single-copy atomic, so no fence and no separate flag are needed.
// Stamps are monotonic per rank; slots are double-buffered by stamp parity (a rank cannot run two
// reductions ahead of a peer, because the next reduction needs the peer's contribution to this one).
// A wait that exceeds spin_limit cycles marks the solve as FS_ERR_COMM instead of hanging the GPU.
#pragma once
#include "fs_context.hpp"

namespace fs {

constexpr int PEER_MAX = 8;                       // ranks of one NVSwitch box
constexpr int MBOX_RED = 0;                       // [2 parities][PEER_MAX ranks][4 values][2 words] LL words
constexpr int MBOX_HALO = 2 * PEER_MAX * 8;       // [PEER_MAX] halo stamps
constexpr int MBOX_WORDS = 256;                   // 2 KB header in front of p

struct PeerWin;
__device__ __forceinline__ unsigned long long global_ns();
struct PeerWin {
    int rank, world;
    unsigned long long seq_red;                   // stamp of this rank's latest reduction contribution
    unsigned long long seq_halo;                  // stamp of this rank's latest halo push
    unsigned long long *mbox[PEER_MAX];           // mailbox of every rank (mbox[rank] = own, local memory)
    double *peer_p[PEER_MAX];                     // p vector of every rank in ITS local layout
    int n_recv, recv_rank[PEER_MAX];              // ranks this rank receives halo values from
    int n_send, send_rank[PEER_MAX];
    long long spin_limit;                         // clock64 ticks
    // measurement (block 0 only): SM cycles spent waiting for the neighbours' halo stamps / the partial sums of the
    // p.Ap reduction / of the r.z reduction, and how many waits were counted (fs_get_comm_stats)
    unsigned long long wait_cycles[3], wait_count[3];
    // wall-clock (globaltimer, ns) from block 0's entry to the end of the last block, summed per kernel:
    // 0 k_spmv_sell, 1 k_update, 2 k_direction; kern_t0 = entry time of the running kernel
    unsigned long long kern_ns[3], kern_t0[3];
};

__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void peer_kern_begin(PeerWin *pw, int k)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) *reinterpret_cast<volatile unsigned long long *>(&pw->kern_t0[k]) = global_ns();
}
// called by ONE thread of the block that finishes last
__device__ __forceinline__ void peer_kern_end(PeerWin *pw, int k)
{
    pw->kern_ns[k] += global_ns() - *reinterpret_cast<volatile unsigned long long *>(&pw->kern_t0[k]);
}

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ double ld_volatile_f64(const double *p)
{
    double v;
    asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_f64(double *p, double v)
{
    asm volatile("st.volatile.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}

// spin until *flag >= stamp; false on timeout
__device__ __forceinline__ bool peer_spin(const unsigned long long *flag, unsigned long long stamp, long long limit)
{
    if (ld_acquire_sys(flag) >= stamp) return true;
    const long long t0 = clock64();
    while (ld_acquire_sys(flag) < stamp) {
        __nanosleep(64);
        if (clock64() - t0 > limit) return false;
    }
    return true;
}

__device__ __forceinline__ void st_relaxed_sys_v2(unsigned long long *p, unsigned long long a, unsigned long long b)
{
    asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ void ld_relaxed_sys_v2(const unsigned long long *p, unsigned long long &a, unsigned long long &b)
{
    asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}

// one thread: this rank's NV partial sums -> every rank's mailbox as LL words
template <int NV>
__device__ __forceinline__ void peer_red_push(PeerWin *pw, const double (&v)[NV])
{
    const unsigned long long stamp = pw->seq_red + 1;
    const unsigned long long tag = (stamp & 0xffffffffull) << 32;
    const int par = (int)(stamp & 1), me = pw->rank;
    unsigned long long w[NV][2];
#pragma unroll
    for (int k = 0; k < NV; k++) {
        const unsigned long long bits = (unsigned long long)__double_as_longlong(v[k]);
        w[k][0] = (bits & 0xffffffffull) | tag;
        w[k][1] = (bits >> 32) | tag;
    }
    // all mailbox pointers first (independent loads), then the stores back to back: a load between two stores would
    // wait for the store's "memory" clobber and add an L2 round trip per rank to the tail of the producing kernel
    unsigned long long *mb[PEER_MAX];
    const int world = pw->world;
#pragma unroll
    for (int r = 0; r < PEER_MAX; r++) mb[r] = pw->mbox[r];
#pragma unroll
    for (int r = 0; r < PEER_MAX; r++)
        if (r < world) {
            unsigned long long *slot = mb[r] + MBOX_RED + ((par * PEER_MAX + me) * 4) * 2;
#pragma unroll
            for (int k = 0; k < NV; k++) st_relaxed_sys_v2(slot + 2 * k, w[k][0], w[k][1]);
        }
    pw->seq_red = stamp;
}

// whole block: wait for every rank's contribution to the reduction this rank produced last, add in rank
// order.  Lane (r, k) of warp 0 polls value k of rank r, so the wait costs one round trip, not `world`.
// Returns false on timeout (state marked FS_ERR_COMM by the caller).
template <int NV>
__device__ __forceinline__ bool peer_red_wait(PeerWin *pw, double (&out)[NV])
{
    __shared__ double s_val[PEER_MAX * 4];
    __shared__ int s_bad;
    if (threadIdx.x == 0) s_bad = 0;
    __syncthreads();
    const int world = pw->world;
    const long long t_begin = (blockIdx.x == 0 && threadIdx.x == 0) ? clock64() : 0;
    if (threadIdx.x < world * NV) {
        const int r = threadIdx.x / NV, k = threadIdx.x - NV * r;
        const unsigned long long stamp = *reinterpret_cast<volatile unsigned long long *>(&pw->seq_red);
        const unsigned long long tag = stamp & 0xffffffffull;
        const unsigned long long *slot = pw->mbox[pw->rank] + MBOX_RED + (((int)(stamp & 1) * PEER_MAX + r) * 4 + k) * 2;
        unsigned long long a, b;
        ld_relaxed_sys_v2(slot, a, b);
        if ((a >> 32) != tag || (b >> 32) != tag) {
            const long long t0 = clock64();
            for (;;) {
                ld_relaxed_sys_v2(slot, a, b);
                if ((a >> 32) == tag && (b >> 32) == tag) break;
                if (clock64() - t0 > pw->spin_limit) { s_bad = 1; break; }
            }
        }
        s_val[4 * r + k] = __longlong_as_double((long long)((a & 0xffffffffull) | (b << 32)));
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NV; k++) {
        double acc = 0.0;
        for (int r = 0; r < world; r++) acc += s_val[4 * r + k];
        out[k] = acc;
    }
    const bool ok = s_bad == 0;
    if (blockIdx.x == 0 && threadIdx.x == 0 && NV <= 2) {   // NV = 1: p.Ap, NV = 2: r.z and the norm
        pw->wait_cycles[NV] += (unsigned long long)(clock64() - t_begin);
        pw->wait_count[NV] += 1;
    }
    __syncthreads();
    return ok;
}

// whole block: wait until every neighbour has pushed as many halos as this rank has
__device__ __forceinline__ bool peer_halo_wait(PeerWin *pw)
{
    __shared__ int s_hok;
    if (threadIdx.x == 0) {
        const long long t_begin = clock64();
        const unsigned long long stamp = *reinterpret_cast<volatile unsigned long long *>(&pw->seq_halo);
        const unsigned long long *mb = pw->mbox[pw->rank];
        bool ok = true;
        for (int k = 0; k < pw->n_recv && ok; k++) ok = peer_spin(mb + MBOX_HALO + pw->recv_rank[k], stamp, pw->spin_limit);
        s_hok = ok ? 1 : 0;
        if (blockIdx.x == 0) {
            pw->wait_cycles[0] += (unsigned long long)(clock64() - t_begin);
            pw->wait_count[0] += 1;
        }
    }
    __syncthreads();
    return s_hok != 0;
}

// one warp: lane 0 waits for the stamps (acquire), the shuffle orders the other lanes' later loads behind it
__device__ __forceinline__ bool peer_halo_wait_warp(PeerWin *pw, int lane)
{
    int ok = 1;
    if (lane == 0) {
        const long long t_begin = clock64();
        const unsigned long long stamp = *reinterpret_cast<volatile unsigned long long *>(&pw->seq_halo);
        const unsigned long long *mb = pw->mbox[pw->rank];
        for (int k = 0; k < pw->n_recv && ok; k++) ok = peer_spin(mb + MBOX_HALO + pw->recv_rank[k], stamp, pw->spin_limit) ? 1 : 0;
        atomicAdd(&pw->wait_cycles[0], (unsigned long long)(clock64() - t_begin));   // every waiting warp counts: average per wait
        atomicAdd(&pw->wait_count[0], 1ull);
    }
    return __shfl_sync(0xffffffffu, ok, 0) != 0;
}

__device__ __forceinline__ void peer_fail(CgState *s)
{
    s->status = FS_ERR_COMM;
    s->done = 1;
}

// The push folded into the kernel that PRODUCES the vector (k_direction): thread t < 3 n_send forms its 16-byte
// piece with `make` and stores it into the neighbour's halo segment; the last of the n_push_blocks leading blocks
// stamps.  The stamp therefore leaves long before the consumer (the neighbour's next SpMV) starts.
template <class Make>
__device__ __forceinline__ void peer_push_inline(PeerWin *pw, int64_t n_send, const int32_t *__restrict__ idx,
                                                 const int32_t *__restrict__ push_peer, const int32_t *__restrict__ push_dst,
                                                 unsigned int *push_counter, int n_push_blocks, Make make)
{
    if ((int)blockIdx.x >= n_push_blocks) return;
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t < 3 * n_send) {
        const int64_t s = t / 3;
        const int h = (int)(t - 3 * s);
        const double2 v = make(idx[s], h);
        double2 *dst = reinterpret_cast<double2 *>(pw->peer_p[push_peer[s]] + 6 * (size_t)push_dst[s]) + h;
        *dst = v;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int ticket = atomicInc(push_counter, (unsigned int)n_push_blocks - 1);
        if (ticket == (unsigned int)n_push_blocks - 1) {
            __threadfence_system();
            const unsigned long long stamp = pw->seq_halo + 1;
            for (int k = 0; k < pw->n_send; k++) st_release_sys(pw->mbox[pw->send_rank[k]] + MBOX_HALO + pw->rank, stamp);
            pw->seq_halo = stamp;
        }
    }
}

// owned boundary values of p -> the neighbours' halo segments (remote stores), then one stamp per neighbour.
// push_peer[s] / push_dst[s]: destination rank and LOCAL node index there of send-list entry s.
static __global__ void __launch_bounds__(256)
k_halo_push(PeerWin *pw, int64_t n_send, const int32_t *__restrict__ idx, const int32_t *__restrict__ push_peer,
            const int32_t *__restrict__ push_dst, const double *__restrict__ vec, CgState *state, unsigned int *counter)
{
    if (state->done) return;
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t < 3 * n_send) {
        const int64_t s = t / 3;
        const int h = (int)(t - 3 * s);
        const double2 v = reinterpret_cast<const double2 *>(vec + 6 * (size_t)idx[s])[h];
        double2 *dst = reinterpret_cast<double2 *>(pw->peer_p[push_peer[s]] + 6 * (size_t)push_dst[s]) + h;
        *dst = v;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int ticket = atomicInc(counter, gridDim.x - 1);
        if (ticket == gridDim.x - 1) {
            __threadfence_system();
            const unsigned long long stamp = pw->seq_halo + 1;
            for (int k = 0; k < pw->n_send; k++) st_release_sys(pw->mbox[pw->send_rank[k]] + MBOX_HALO + pw->rank, stamp);
            pw->seq_halo = stamp;
        }
    }
}

}  // namespace fs

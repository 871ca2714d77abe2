// fs_peer.cu -- host side of the NVLink peer window (device side: fs_peer.cuh).
//
// Every rank allocates [1 KB mailbox | p vector], exports it with cudaIpcGetMemHandle, all-gathers the
// handles (plus the local index at which each neighbour's values land in its halo) over the NCCL
// communicator that fs_dist_init created, and maps the other ranks' windows.  Replaces, for the Krylov
// loop, PETSc's VecScatter + MPI_Allreduce (fs.cpp:138).  Collective: called from fs_set_mesh on every rank.
#include <algorithm>
#include <cstring>

#include "fs_context.hpp"
#include "fs_nccl.hpp"
#include "fs_peer.cuh"

namespace fs {

namespace {
struct PeerMeta {                     // what one rank tells the others (fixed 256 bytes)
    cudaIpcMemHandle_t handle;        // 64 bytes
    int64_t n_local;
    int64_t recv_off_from[PEER_MAX];  // LOCAL node index where rank r's values start in my halo (-1: none)
    int32_t ok;
    char pad[256 - 64 - 8 - 8 * PEER_MAX - 4];
};
static_assert(sizeof(PeerMeta) == 256, "PeerMeta layout");

int nccl_barrier(fs_context *c, int *all_ok)
{
    DevBuf<int> flag;
    FS_CUDA(c, flag.alloc(1));
    int v = all_ok ? *all_ok : 1;
    FS_CUDA(c, cudaMemcpyAsync(flag.p, &v, sizeof v, cudaMemcpyHostToDevice, c->stream));
    ncclResult_t r = nccl().AllReduce(flag.p, flag.p, 1, ncclInt, ncclMin, (ncclComm_t)c->comm, c->stream);
    if (r != ncclSuccess) return fail(c, FS_ERR_COMM, std::string("ncclAllReduce: ") + nccl().GetErrorString(r));
    FS_CUDA(c, cudaMemcpyAsync(&v, flag.p, sizeof v, cudaMemcpyDeviceToHost, c->stream));
    FS_CUDA(c, cudaStreamSynchronize(c->stream));
    if (all_ok) *all_ok = v;
    return FS_OK;
}
}  // namespace

int peer_window_teardown(fs_context *c)
{
    if (!c->win_base) return FS_OK;
    cudaSetDevice(c->device);
    PhaseTimer tm("peer_window_teardown");
    if (c->stream) cudaStreamSynchronize(c->stream);
    // nobody may still be pushing into a window that is about to disappear
    if (c->comm && c->world > 1) nccl_barrier(c, nullptr);
    tm.lap("barrier");
    for (int r = 0; r < PEER_MAX; r++)
        if (c->peer_base[r]) {
            cudaIpcCloseMemHandle(c->peer_base[r]);
            c->peer_base[r] = nullptr;
        }
    tm.lap("cudaIpcCloseMemHandle");
    if (c->comm && c->world > 1) nccl_barrier(c, nullptr);
    c->d_p.release();  // view into the window
    cudaFree(c->win_base);
    tm.lap("barrier + cudaFree");
    c->win_base = nullptr;
    c->peer_ready = false;
    return FS_OK;
}

int peer_window_setup(fs_context *c)
{
    c->peer_ready = false;
    if (c->world <= 1 || c->comm_pref == FS_COMM_NCCL) return FS_OK;
    if (c->world > PEER_MAX) {
        if (c->comm_pref == FS_COMM_PEER) return fail(c, FS_ERR_ARG, "peer windows support at most 8 ranks");
        return FS_OK;
    }
    cudaStream_t st = c->stream;
    const int me = c->rank, W = c->world;
    int ok = 1;
    PhaseTimer tm("peer_window_setup");

    // ---- own window: [mailbox | p]; p moves into it ----
    const size_t p_bytes = sizeof(double) * 6 * (size_t)c->n_local;
    void *base = nullptr;
    if (cudaMalloc(&base, MBOX_WORDS * 8 + p_bytes) != cudaSuccess) { cudaGetLastError(); ok = 0; }
    PeerMeta mine;
    memset(&mine, 0, sizeof mine);
    if (ok) {
        FS_CUDA(c, cudaMemsetAsync(base, 0, MBOX_WORDS * 8 + p_bytes, st));
        if (cudaIpcGetMemHandle(&mine.handle, base) != cudaSuccess) { cudaGetLastError(); ok = 0; }
    }
    mine.n_local = c->n_local;
    for (int r = 0; r < PEER_MAX; r++) mine.recv_off_from[r] = -1;
    for (const Peer &pr : c->peers)
        if (pr.recv_count > 0 && pr.rank < PEER_MAX) mine.recv_off_from[pr.rank] = pr.recv_off;
    mine.ok = ok;
    tm.lap("window alloc + export");

    // ---- all-gather the metadata ----
    DevBuf<char> d_meta;
    FS_CUDA(c, d_meta.alloc(sizeof(PeerMeta) * (size_t)(W + 1)));
    FS_CUDA(c, cudaMemcpyAsync(d_meta.p + sizeof(PeerMeta) * W, &mine, sizeof mine, cudaMemcpyHostToDevice, st));
    ncclResult_t nr = nccl().AllGather(d_meta.p + sizeof(PeerMeta) * W, d_meta.p, sizeof(PeerMeta), ncclChar, (ncclComm_t)c->comm, st);
    if (nr != ncclSuccess) {
        if (base) cudaFree(base);
        return fail(c, FS_ERR_COMM, std::string("ncclAllGather: ") + nccl().GetErrorString(nr));
    }
    std::vector<PeerMeta> meta(W);
    FS_CUDA(c, cudaMemcpyAsync(meta.data(), d_meta.p, sizeof(PeerMeta) * W, cudaMemcpyDeviceToHost, st));
    FS_CUDA(c, cudaStreamSynchronize(st));
    for (int r = 0; r < W; r++) ok = ok && meta[r].ok;
    tm.lap("all-gather of the handles");

    // ---- map the other windows ----
    void *pb[PEER_MAX] = {};
    if (ok)
        for (int r = 0; r < W && ok; r++) {
            if (r == me) continue;
            if (cudaIpcOpenMemHandle(&pb[r], meta[r].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                pb[r] = nullptr;
                ok = 0;
            }
        }
    tm.lap("cudaIpcOpenMemHandle");
    // ---- device-side description: everything that can fail locally happens BEFORE the one collective decision ----
    PeerWin h;
    memset(&h, 0, sizeof h);
    h.rank = me;
    h.world = W;
    h.spin_limit = 20000000000LL;  // ~10 s of SM clocks
    std::string why;
    std::vector<int32_t> push_peer(c->send_total), push_dst(c->send_total);
    if (ok) {
        for (int r = 0; r < W; r++) {
            char *b = (char *)(r == me ? base : pb[r]);
            h.mbox[r] = (unsigned long long *)b;
            h.peer_p[r] = (double *)(b + MBOX_WORDS * 8);
        }
        for (const Peer &pr : c->peers) {
            if (pr.recv_count > 0) h.recv_rank[h.n_recv++] = pr.rank;
            if (pr.send_count > 0) h.send_rank[h.n_send++] = pr.rank;
        }
        for (const Peer &pr : c->peers) {
            const int64_t off = meta[pr.rank].recv_off_from[me];
            if (pr.send_count > 0 && off < 0) {  // the neighbour does not expect values from this rank
                ok = 0;
                why = "halo plans of neighbouring ranks disagree";
                break;
            }
            for (int64_t k = 0; k < pr.send_count; k++) {
                push_peer[pr.send_off + k] = pr.rank;
                push_dst[pr.send_off + k] = (int32_t)(off + k);
            }
        }
    }
    std::vector<uint8_t> is_send((size_t)std::max<int64_t>(c->n_own, 1), 0);
    c->push_foldable = true;
    if (ok) {  // send_idx holds LOCAL node ids of owned nodes; a node listed for two neighbours cannot use the folded push
        std::vector<int32_t> sidx(c->send_total);
        if (c->send_total && cudaMemcpy(sidx.data(), c->d_send_idx.p, sizeof(int32_t) * c->send_total, cudaMemcpyDeviceToHost) != cudaSuccess) {
            cudaGetLastError();
            ok = 0;
            why = "reading the send list failed";
        }
        for (int64_t k = 0; ok && k < c->send_total; k++) {
            const int64_t o = sidx[k] - c->own_lo;
            if (o < 0 || o >= c->n_own) { ok = 0; why = "send list holds a node this rank does not own"; break; }
            if (is_send[o]) c->push_foldable = false;
            is_send[o] = 1;
        }
    }
    if (ok) {
        bool up = c->d_is_send.alloc(is_send.size()) == cudaSuccess &&
                  cudaMemcpy(c->d_is_send.p, is_send.data(), is_send.size(), cudaMemcpyHostToDevice) == cudaSuccess;
        if (!up) {
            cudaGetLastError();
            ok = 0;
            why = "device allocation of the peer tables failed";
        }
    }
    if (ok) {
        bool up = c->d_pw.alloc(1) == cudaSuccess && c->d_push_peer.alloc(push_peer.size()) == cudaSuccess &&
                  c->d_push_dst.alloc(push_dst.size()) == cudaSuccess &&
                  cudaMemcpy(c->d_pw.p, &h, sizeof h, cudaMemcpyHostToDevice) == cudaSuccess;
        if (up && !push_peer.empty())
            up = cudaMemcpy(c->d_push_peer.p, push_peer.data(), sizeof(int32_t) * push_peer.size(), cudaMemcpyHostToDevice) == cudaSuccess &&
                 cudaMemcpy(c->d_push_dst.p, push_dst.data(), sizeof(int32_t) * push_dst.size(), cudaMemcpyHostToDevice) == cudaSuccess;
        if (!up) {
            cudaGetLastError();
            ok = 0;
            why = "device allocation of the peer tables failed";
        }
    }
    tm.lap("tables");
    // folding the halo push into k_direction changes WHEN a rank pushes (and how often per solve): all ranks or none
    {
        const int64_t vg = std::max<int64_t>(1, std::min<int64_t>((c->n_own + 255) / 256, (int64_t)c->sm_count * c->vec_blocks_per_sm));
        int fold = (c->push_foldable && 3 * c->send_total <= vg * 256) ? 1 : 0;
        int rcf = nccl_barrier(c, &fold);
        if (rcf != FS_OK) ok = 0;
        c->push_foldable = fold != 0;
    }
    // every rank must take the same path; the all-reduce is also the barrier "all mailboxes are zeroed"
    int all_ok = ok;
    int rc = nccl_barrier(c, &all_ok);
    if (rc != FS_OK || !all_ok) {
        for (int r = 0; r < W; r++)
            if (pb[r]) cudaIpcCloseMemHandle(pb[r]);
        if (rc == FS_OK) rc = nccl_barrier(c, nullptr);  // nobody frees a window another rank still has mapped
        if (base) cudaFree(base);
        c->d_pw.release();
        c->d_push_peer.release();
        c->d_push_dst.release();
        if (rc != FS_OK) return rc;
        if (!why.empty()) return fail(c, FS_ERR_STATE, "peer windows: " + why);
        if (c->comm_pref == FS_COMM_PEER) return fail(c, FS_ERR_COMM, "peer windows unavailable (cudaIpc / P2P mapping or set-up failed on some rank)");
        return FS_OK;  // AUTO: stay on the NCCL path
    }

    tm.lap("two verdict barriers");
    c->win_base = base;
    for (int r = 0; r < W; r++) c->peer_base[r] = pb[r];
    c->d_p.view((double *)((char *)base + MBOX_WORDS * 8), 6 * (size_t)c->n_local);
    c->peer_ready = true;
    return FS_OK;
}

}  // namespace fs

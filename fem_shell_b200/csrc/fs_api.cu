// fs_api.cu -- the C ABI of include/femshell_b200.h: context life cycle, mesh ingestion (DOF map,
// Dirichlet sets, node-block partition, halo lists), loads, solution gather, exports.
// Reference anchors are given per function in the header; host-side restatements here follow
//   DOF numbering      libMesh DofMap as used at fs.cpp:125,1205
//   Dirichlet sets     fs.cpp:90-120
//   interface nodes    fsp.cpp:55-71
//   solution layout    fs.cpp:140-141,163-169
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstring>
#include <numeric>

#include "fs_context.hpp"
#include "fs_nccl.hpp"
#include "fs_partition.hpp"
#include "fs_peer.cuh"
#include "fs_host_par.hpp"

using namespace fs;

#define FS_TRY(expr)               \
    do {                           \
        int rc__ = (expr);         \
        if (rc__ != FS_OK) return rc__; \
    } while (0)

#define FS_CHECK_CTX(c) \
    if (!(c)) return FS_ERR_ARG

static inline unsigned int nblk(int64_t n, int bs) { return (unsigned int)((n + bs - 1) / bs); }

extern "C" {

int fs_create(fs_context **out, int device)
{
    if (!out) return FS_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) return FS_ERR_CUDA;  // no CPU fallback
    fs_context *c = new (std::nothrow) fs_context();
    if (!c) return FS_ERR_ARG;
    c->device = device;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&c->ev0) != cudaSuccess || cudaEventCreate(&c->ev1) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_copy, cudaEventDisableTiming) != cudaSuccess ||
        cudaMallocHost((void **)&c->h_state, sizeof(CgState)) != cudaSuccess) {
        delete c;
        return FS_ERR_CUDA;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) c->sm_count = prop.multiProcessorCount;
    if (c->d_state.alloc(1) != cudaSuccess || c->d_counter.alloc(4) != cudaSuccess || c->d_flag.alloc(1) != cudaSuccess) {
        delete c;
        return FS_ERR_CUDA;
    }
    solver_query_occupancy(c);
    *out = c;
    return FS_OK;
}

int fs_destroy(fs_context *c)
{
    FS_CHECK_CTX(c);
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->cg_graph_exec) cudaGraphExecDestroy(c->cg_graph_exec);
    peer_window_teardown(c);
    if (c->comm && nccl().ok) nccl().CommDestroy((ncclComm_t)c->comm);
    if (c->h_state) cudaFreeHost(c->h_state);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->ev_copy) cudaEventDestroy(c->ev_copy);
    for (cudaEvent_t e : c->prof.ev)
        if (e) cudaEventDestroy(e);
    if (c->copy_stream) {
        cudaStreamSynchronize(c->copy_stream);
        cudaStreamDestroy(c->copy_stream);
    }
    cudaStream_t st = c->stream;
    delete c;  // frees the device buffers
    if (st) cudaStreamDestroy(st);
    return FS_OK;
}

const char *fs_last_error(const fs_context *c) { return c ? c->err.c_str() : "null context"; }

void *fs_get_stream(fs_context *c) { return c ? (void *)c->stream : nullptr; }

int fs_dist_unique_id(uint8_t id_out[128])
{
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    ncclUniqueId id;
    if (!nccl().ok || nccl().GetUniqueId(&id) != ncclSuccess) return FS_ERR_COMM;
    memcpy(id_out, &id, 128);
    return FS_OK;
}

int fs_dist_init(fs_context *c, int rank, int world, const uint8_t id_bytes[128])
{
    FS_CHECK_CTX(c);
    if (world < 1 || rank < 0 || rank >= world) return fail(c, FS_ERR_ARG, "bad rank/world");
    if (c->n_nodes) return fail(c, FS_ERR_STATE, "fs_dist_init must precede fs_set_mesh");
    c->rank = rank;
    c->world = world;
    if (world > 1) {
        FS_CUDA(c, cudaSetDevice(c->device));
        ncclUniqueId id;
        memcpy(&id, id_bytes, 128);
        ncclComm_t comm;
        if (!nccl().ok) return fail(c, FS_ERR_COMM, "libnccl.so.2 could not be loaded");
        // this library only sends halo slabs point to point and all-reduces a few doubles: NVLink SHARP brings nothing
        // here, and its channel set-up costs about a second at the first collective.  The user's setting wins.
        setenv("NCCL_NVLS_ENABLE", "0", 0);
        ncclResult_t r = nccl().CommInitRank(&comm, world, id, rank);
        if (r != ncclSuccess) return fail(c, FS_ERR_COMM, std::string("ncclCommInitRank: ") + nccl().GetErrorString(r));
        c->comm = (ncclComm *)comm;
        // NCCL connects its channels at the FIRST collective of each kind (0.3 s on 2 GPUs, 1.1 s on 8: measured as the
        // "peer window" phase of the first fs_set_mesh, profiles/r02h).  That is communicator set-up: pay it here, once,
        // with the three operations the library uses (all-reduce, all-gather, send/recv with the neighbouring ranks).
        PhaseTimer tm("fs_dist_init");
        DevBuf<double> w;
        FS_CUDA(c, w.alloc(64 * (size_t)(world + 1)));
        FS_CUDA(c, cudaMemsetAsync(w.p, 0, sizeof(double) * 64 * (world + 1), c->stream));
        r = nccl().AllReduce(w.p, w.p, 4, ncclDouble, ncclSum, comm, c->stream);
        if (r == ncclSuccess) r = nccl().AllGather(w.p + 64 * world, w.p, 256, ncclChar, comm, c->stream);
        if (r == ncclSuccess) r = nccl().GroupStart();
        for (int nb = rank - 1; nb <= rank + 1 && r == ncclSuccess; nb += 2)
            if (nb >= 0 && nb < world) {
                r = nccl().Send(w.p, 8, ncclDouble, nb, comm, c->stream);
                if (r == ncclSuccess) r = nccl().Recv(w.p + 8 + 8 * (nb > rank), 8, ncclDouble, nb, comm, c->stream);
            }
        if (r == ncclSuccess) r = nccl().GroupEnd();
        if (r != ncclSuccess) return fail(c, FS_ERR_COMM, std::string("communicator warm-up: ") + nccl().GetErrorString(r));
        FS_CUDA(c, cudaStreamSynchronize(c->stream));
        tm.lap("first collectives (channel set-up)");
    }
    return FS_OK;
}

int fs_set_comm_mode(fs_context *c, int mode)
{
    FS_CHECK_CTX(c);
    if (mode != FS_COMM_AUTO && mode != FS_COMM_NCCL && mode != FS_COMM_PEER) return fail(c, FS_ERR_ARG, "unknown comm mode");
    if (c->n_nodes) return fail(c, FS_ERR_STATE, "fs_set_comm_mode must precede fs_set_mesh");
    c->comm_pref = mode;
    return FS_OK;
}

int fs_get_comm_mode(fs_context *c, int *mode)
{
    FS_CHECK_CTX(c);
    if (!mode) return fail(c, FS_ERR_ARG, "null output");
    *mode = c->peer_ready ? FS_COMM_PEER : FS_COMM_NCCL;
    return FS_OK;
}

int fs_get_comm_stats(fs_context *c, double out[9], int reset)
{
    FS_CHECK_CTX(c);
    if (!out) return fail(c, FS_ERR_ARG, "null output");
    for (int k = 0; k < 9; k++) out[k] = 0.0;
    if (!c->peer_ready) return FS_OK;
    FS_CUDA(c, cudaSetDevice(c->device));
    FS_CUDA(c, cudaStreamSynchronize(c->stream));
    PeerWin h;
    FS_CUDA(c, cudaMemcpy(&h, c->d_pw.p, sizeof h, cudaMemcpyDeviceToHost));
    int khz = 0;
    FS_CUDA(c, cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, c->device));   // clock64 ticks at the SM clock
    for (int k = 0; k < 3; k++) {
        out[k] = khz > 0 ? 1e3 * (double)h.wait_cycles[k] / (double)khz : 0.0;
        out[3 + k] = (double)h.wait_count[k];
        out[6 + k] = 1e-3 * (double)h.kern_ns[k];
    }
    if (reset) {
        const size_t off = offsetof(PeerWin, wait_cycles);
        FS_CUDA(c, cudaMemset((char *)c->d_pw.p + off, 0, sizeof h.wait_cycles + sizeof h.wait_count + sizeof h.kern_ns + sizeof h.kern_t0));
    }
    return FS_OK;
}

int fs_set_material(fs_context *c, double nu, double E, double t)
{
    FS_CHECK_CTX(c);
    if (!(E > 0.0) || !(t > 0.0) || !(nu > -1.0 && nu < 1.0)) return fail(c, FS_ERR_ARG, "material out of range");
    c->nu = nu; c->E = E; c->thickness = t;
    c->material_set = true;
    c->assembled = false;
    return FS_OK;
}

int fs_set_quirks(fs_context *c, int q)
{
    FS_CHECK_CTX(c);
    c->quirks = q & FS_QUIRKS_REFERENCE;
    c->assembled = false;
    return FS_OK;
}

int fs_set_dof_order(fs_context *c, int mode)
{
    FS_CHECK_CTX(c);
    if (mode != FS_DOF_FIRST_ENCOUNTER && mode != FS_DOF_NODE_ID) return fail(c, FS_ERR_ARG, "unknown dof order");
    if (c->n_nodes) return fail(c, FS_ERR_STATE, "fs_set_dof_order must precede fs_set_mesh");
    c->dof_mode = mode;
    return FS_OK;
}

int fs_set_spmv_format(fs_context *c, int mode)
{
    FS_CHECK_CTX(c);
    if (mode != FS_SPMV_AUTO && mode != FS_SPMV_FULL) return fail(c, FS_ERR_ARG, "unknown SpMV format");
    if (mode != c->spmv_format_pref) c->sell_checked = c->sell_active = false;
    c->spmv_format_pref = mode;
    return FS_OK;
}

int fs_get_spmv_format(fs_context *c, int64_t info[4])
{
    FS_CHECK_CTX(c);
    if (!c->assembled || !info) return fail(c, FS_ERR_STATE, "not assembled / null output");
    FS_CUDA(c, cudaSetDevice(c->device));
    FS_TRY(spmv_format_prepare(c));
    if (c->sell_active) {
        info[0] = c->sell_nz;
        info[1] = 8 * (int64_t)c->sell_nz * 32 * c->sell_slots + 4 * 32 * c->sell_slots + 4 * (c->sell_slices + 1);
        info[2] = 32 * c->sell_slots;
    } else {
        info[0] = 36;
        info[1] = 8 * 36 * c->n_blocks + 4 * c->n_blocks + 4 * (c->n_own + 1);
        info[2] = c->n_blocks;
    }
    info[3] = (int64_t)c->sell_detected;
    return FS_OK;
}

int fs_set_assembly_mode(fs_context *c, int mode)
{
    FS_CHECK_CTX(c);
    if (mode != FS_ASM_COLORED && mode != FS_ASM_GATHER) return fail(c, FS_ERR_ARG, "unknown assembly mode");
    c->asm_mode = mode;
    return FS_OK;
}

// ---------------------------------------------------------------------------------------------
// mesh ingestion
// ---------------------------------------------------------------------------------------------
int fs_set_mesh(fs_context *c, int64_t n_nodes, const double *xyz, int64_t n_elem, const int32_t *etype,
                const int64_t *eptr, const int32_t *enodes, int64_t n_bc, const int32_t *bc)
{
    FS_CHECK_CTX(c);
    if (n_nodes <= 0 || n_elem <= 0 || !xyz || !etype || !eptr || !enodes || (n_bc > 0 && !bc))
        return fail(c, FS_ERR_ARG, "empty mesh or null array");
    if (n_nodes >= (int64_t)1 << 31 || n_elem >= (int64_t)1 << 31) return fail(c, FS_ERR_ARG, "mesh too large for 32-bit ids");
    FS_CUDA(c, cudaSetDevice(c->device));
    PhaseTimer tm("fs_set_mesh");
    // every argument check comes BEFORE the first collective step: a bad mesh on one rank must not leave the others
    // waiting in a barrier
    const int ht = host_threads(c->world);
    {
        std::vector<int64_t> bad_type(ht, -1), bad_node(ht, -1);
        parallel_chunks(n_elem, ht, [&](int t, int64_t e0, int64_t e1) {
            for (int64_t e = e0; e < e1; e++) {
                const int nen = (int)(eptr[e + 1] - eptr[e]);
                if ((etype[e] == FS_TRI3 && nen != 3) || (etype[e] == FS_QUAD4 && nen != 4) || (etype[e] != FS_TRI3 && etype[e] != FS_QUAD4)) {
                    if (bad_type[t] < 0) bad_type[t] = e;
                    continue;
                }
                for (int64_t k = eptr[e]; k < eptr[e + 1]; k++)
                    if ((enodes[k] < 0 || enodes[k] >= n_nodes) && bad_node[t] < 0) bad_node[t] = e;
            }
        });
        for (int t = 0; t < ht; t++) {
            if (bad_type[t] >= 0) return fail(c, FS_ERR_ARG, "element " + std::to_string(bad_type[t]) + ": only TRI3 (3) and QUAD4 (5) are supported");
            if (bad_node[t] >= 0) return fail(c, FS_ERR_ARG, "node id out of range in element " + std::to_string(bad_node[t]));
        }
    }
    for (int64_t i = 0; i < n_bc; i++) {
        const int32_t e = bc[3 * i], sd = bc[3 * i + 1];
        if (e < 0 || e >= n_elem) return fail(c, FS_ERR_ARG, "boundary record " + std::to_string(i) + ": element out of range");
        if (sd < 0 || sd >= (int)(eptr[e + 1] - eptr[e])) return fail(c, FS_ERR_ARG, "boundary record " + std::to_string(i) + ": side out of range");
    }
    FS_TRY(peer_window_teardown(c));
    c->n_nodes = n_nodes;
    c->n_elem = n_elem;
    c->pattern_ready = c->assembled = c->loads_set = c->rhs_ready = c->have_solution = false;
    c->gather_ready = c->gather_unavailable = false;
    c->slice_ready = c->parity_valid = false;
    c->sell_checked = c->sell_active = c->sell_layout_ready = false;
    c->ml_geom_ready = c->ml_values_ready = false;
    if (c->cg_graph_exec) { cudaGraphExecDestroy(c->cg_graph_exec); c->cg_graph_exec = nullptr; }

    tm.lap("validate + teardown");
    // ---- DOF order (a12) ----
    const int64_t n_g = compute_dof_order(c->dof_mode, n_nodes, n_elem, eptr, enodes, c->dofnode);
    c->n_dofnodes_global = n_g;
    c->node_of_dof.resize(n_g);
    parallel_chunks(n_nodes, ht, [&](int, int64_t i0, int64_t i1) {
        for (int64_t i = i0; i < i1; i++)
            if (c->dofnode[i] >= 0) c->node_of_dof[c->dofnode[i]] = (int32_t)i;   // a permutation: distinct targets
    });

    tm.lap("dof order");
    // bounding box of the nodes that carry DOFs (planarity test below; lattice of the multilevel preconditioner).  The
    // largest element extent per axis, which sizes the cells of the first lattice, is measured on the device when the
    // hierarchy is first built (fs_mlpc.cu ml_element_extents): here it was a pass of random reads over the whole mesh
    {
        const double inf = 1e300;
        std::vector<double> lo(3 * ht, inf), hi(3 * ht, -inf);
        parallel_chunks(n_nodes, ht, [&](int t, int64_t i0, int64_t i1) {
            double l[3] = {inf, inf, inf}, h[3] = {-inf, -inf, -inf};   // thread-local: the shared arrays are written once
            for (int64_t i = i0; i < i1; i++) {
                if (c->dofnode[i] < 0) continue;
                for (int d = 0; d < 3; d++) {
                    const double v = xyz[3 * i + d];
                    l[d] = std::min(l[d], v);
                    h[d] = std::max(h[d], v);
                }
            }
            for (int d = 0; d < 3; d++) { lo[3 * t + d] = l[d]; hi[3 * t + d] = h[d]; }
        });
        for (int d = 0; d < 3; d++) {
            c->bbox_lo[d] = inf; c->bbox_hi[d] = -inf; c->ml_h[d] = 0.0;
            for (int t = 0; t < ht; t++) {
                c->bbox_lo[d] = std::min(c->bbox_lo[d], lo[3 * t + d]);
                c->bbox_hi[d] = std::max(c->bbox_hi[d], hi[3 * t + d]);
            }
            if (c->bbox_lo[d] > c->bbox_hi[d]) c->bbox_lo[d] = c->bbox_hi[d] = 0.0;
        }
    }
    // ---- planar shell?  (fs_context.hpp: plane frame of the slice pass / compacted SpMV) ----
    c->planar = c->plane_rot = false;
    double plane_ref[3] = {0, 0, 0};
    {
        const double Qi[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        memcpy(c->plane_Q, Qi, sizeof Qi);
        if (c->bbox_hi[2] == c->bbox_lo[2]) c->planar = true;   // the xy plane itself: nothing to rotate
        else {
            // frame of element 0 (fs.cpp:315-390: x^ along the first edge / the mid-side line, n the element normal)
            const int32_t *en = enodes + eptr[0];
            const int nen = (int)(eptr[1] - eptr[0]);
            double A[3], U[3], V[3], n[3];
            for (int d = 0; d < 3; d++) {
                A[d] = xyz[3 * (int64_t)en[0] + d];
                U[d] = xyz[3 * (int64_t)en[1] + d] - A[d];
                V[d] = xyz[3 * (int64_t)en[nen - 1] + d] - A[d];
            }
            n[0] = U[1] * V[2] - U[2] * V[1]; n[1] = U[2] * V[0] - U[0] * V[2]; n[2] = U[0] * V[1] - U[1] * V[0];
            const double lu = std::sqrt(U[0] * U[0] + U[1] * U[1] + U[2] * U[2]), ln = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
            double diam = 0.0;
            for (int d = 0; d < 3; d++) diam += (c->bbox_hi[d] - c->bbox_lo[d]) * (c->bbox_hi[d] - c->bbox_lo[d]);
            diam = std::sqrt(diam);
            if (lu > 0.0 && ln > 0.0 && diam > 0.0) {
                for (int d = 0; d < 3; d++) { U[d] /= lu; n[d] /= ln; }
                const double W[3] = {n[1] * U[2] - n[2] * U[1], n[2] * U[0] - n[0] * U[2], n[0] * U[1] - n[1] * U[0]};   // y^ = n x x^
                std::vector<double> worst(ht, 0.0);
                parallel_chunks(n_nodes, ht, [&](int t, int64_t i0, int64_t i1) {
                    double w = 0.0;
                    for (int64_t i = i0; i < i1; i++) {
                        if (c->dofnode[i] < 0) continue;
                        w = std::max(w, std::fabs(n[0] * (xyz[3 * i] - A[0]) + n[1] * (xyz[3 * i + 1] - A[1]) + n[2] * (xyz[3 * i + 2] - A[2])));
                    }
                    worst[t] = w;
                });
                double w = 0.0;
                for (int t = 0; t < ht; t++) w = std::max(w, worst[t]);
                if (w <= 1e-12 * diam) {
                    c->planar = c->plane_rot = true;
                    for (int d = 0; d < 3; d++) { c->plane_Q[d] = U[d]; c->plane_Q[3 + d] = W[d]; c->plane_Q[6 + d] = n[d]; plane_ref[d] = A[d]; }
                }
            }
        }
    }
    tm.lap("bbox + planarity");
    // ---- Dirichlet bits (fs.cpp:90-120) and coupling interface (fsp.cpp:55-71) ----
    c->node_mask.assign(n_nodes, 0);
    std::vector<uint8_t> is_if(n_nodes, 0);
    for (int64_t i = 0; i < n_bc; i++) {
        const int32_t e = bc[3 * i], s = bc[3 * i + 1], id = bc[3 * i + 2];
        const int nen = (int)(eptr[e + 1] - eptr[e]);
        const int32_t n0 = enodes[eptr[e] + s], n1 = enodes[eptr[e] + (s + 1) % nen];
        uint8_t m = 0;
        if (id == 0 || id == 20) m = 0x07;
        if (id == 1 || id == 21) m = 0x3f;
        c->node_mask[n0] |= m;
        c->node_mask[n1] |= m;
        if (id == 2 || id == 20 || id == 21) is_if[n0] = is_if[n1] = 1;
    }
    c->iface_nodes.clear();
    for (int64_t i = 0; i < n_nodes; i++)
        if (is_if[i]) c->iface_nodes.push_back((int32_t)i);
    c->sols.clear();  // coupled-step state (fsp.cpp:77-87): sized by the first fs_step, not by every mesh
    c->pre_sols.clear();

    tm.lap("dirichlet + interface");
    // ---- node-block partition of the DOF order (fs_partition.cpp) ----
    PartitionPlan plan;
    if (plan_partition(c->dofnode, n_g, n_elem, eptr, enodes, c->rank, c->world, plan, ht) != FS_OK)
        return fail(c, FS_ERR_ARG, "fewer nodes than ranks");
    c->own_begin = plan.own_begin;
    c->own_end = plan.own_end;
    c->n_own = plan.own_end - plan.own_begin;
    c->own_lo = plan.own_lo;
    c->local_to_global = plan.local_to_global;
    c->n_local = (int64_t)c->local_to_global.size();
    std::vector<int32_t> g2l;
    if (c->world > 1) {   // one rank: local == global
        g2l.assign(n_g, -1);
        for (int64_t l = 0; l < c->n_local; l++) g2l[c->local_to_global[l]] = (int32_t)l;
    }
    const std::vector<int32_t> &loc_elems = plan.loc_elems;
    const std::vector<int32_t> &send_idx = plan.send_idx;
    c->peers.clear();
    for (const PeerPlan &pp : plan.peers) {
        Peer pr;
        pr.rank = pp.rank;
        pr.send_count = pp.send_count; pr.send_off = pp.send_off;
        pr.recv_count = pp.recv_count; pr.recv_off = pp.recv_off;
        c->peers.push_back(pr);
    }
    c->send_total = (int64_t)send_idx.size();
    FS_CUDA(c, c->d_send_idx.alloc(send_idx.size()));
    FS_CUDA(c, c->d_sendbuf.alloc(6 * send_idx.size()));
    if (!send_idx.empty())
        FS_CUDA(c, cudaMemcpy(c->d_send_idx.p, send_idx.data(), sizeof(int32_t) * send_idx.size(), cudaMemcpyHostToDevice));

    tm.lap("partition plan");
    // ---- device mesh in local numbering ----
    std::vector<double> lxyz(3 * c->n_local);
    std::vector<uint8_t> lmask(c->n_local);
    parallel_chunks(c->n_local, ht, [&](int, int64_t l0, int64_t l1) {
        for (int64_t l = l0; l < l1; l++) {
            const int32_t node = c->node_of_dof[c->local_to_global[l]];
            lxyz[3 * l + 0] = xyz[3 * (int64_t)node + 0];
            lxyz[3 * l + 1] = xyz[3 * (int64_t)node + 1];
            lxyz[3 * l + 2] = xyz[3 * (int64_t)node + 2];
            lmask[l] = c->node_mask[node];
        }
    });
    FS_CUDA(c, c->d_xyz.alloc(3 * c->n_local));
    FS_CUDA(c, c->d_mask.alloc(c->n_local));
    FS_CUDA(c, cudaMemcpy(c->d_xyz.p, lxyz.data(), sizeof(double) * 3 * c->n_local, cudaMemcpyHostToDevice));
    FS_CUDA(c, cudaMemcpy(c->d_mask.p, lmask.data(), c->n_local, cudaMemcpyHostToDevice));
    c->d_xyz_plane.release();
    if (c->plane_rot) {   // coordinates in the plane frame; the normal component is the same constant for every node
        const double *Q = c->plane_Q;
        const double zc = Q[6] * plane_ref[0] + Q[7] * plane_ref[1] + Q[8] * plane_ref[2];
        std::vector<double> pxyz(3 * c->n_local);
        parallel_chunks(c->n_local, ht, [&](int, int64_t l0, int64_t l1) {
            for (int64_t l = l0; l < l1; l++) {
                const double *x = &lxyz[3 * l];
                pxyz[3 * l + 0] = Q[0] * x[0] + Q[1] * x[1] + Q[2] * x[2];
                pxyz[3 * l + 1] = Q[3] * x[0] + Q[4] * x[1] + Q[5] * x[2];
                pxyz[3 * l + 2] = zc;
            }
        });
        FS_CUDA(c, c->d_xyz_plane.alloc(3 * c->n_local));
        FS_CUDA(c, cudaMemcpy(c->d_xyz_plane.p, pxyz.data(), sizeof(double) * 3 * c->n_local, cudaMemcpyHostToDevice));
    }

    tm.lap("local xyz/mask + upload");
    // local connectivity in local node ids, triangles and quads apart, element order kept: count per chunk, then fill
    std::vector<int32_t> tri, quad, tri_gid, quad_gid;
    {
        const int64_t nle = (int64_t)loc_elems.size();
        std::vector<int64_t> ct(ht + 1, 0), cq(ht + 1, 0);
        const int used = parallel_chunks(nle, ht, [&](int t, int64_t i0, int64_t i1) {
            int64_t a = 0, b = 0;
            for (int64_t i = i0; i < i1; i++) (eptr[loc_elems[i] + 1] - eptr[loc_elems[i]] == 3 ? a : b)++;
            ct[t + 1] = a;
            cq[t + 1] = b;
        });
        for (int t = 0; t < used; t++) { ct[t + 1] += ct[t]; cq[t + 1] += cq[t]; }
        c->n_tri = ct[used];
        c->n_quad = cq[used];
        tri.resize(3 * c->n_tri); tri_gid.resize(c->n_tri);
        quad.resize(4 * c->n_quad); quad_gid.resize(c->n_quad);
        const bool one = c->world == 1;
        parallel_chunks(nle, ht, [&](int t, int64_t i0, int64_t i1) {
            int64_t a = ct[t], b = cq[t];
            for (int64_t i = i0; i < i1; i++) {
                const int32_t e = loc_elems[i];
                const int nen = (int)(eptr[e + 1] - eptr[e]);
                int32_t *dst = nen == 3 ? &tri[3 * a] : &quad[4 * b];
                for (int k = 0; k < nen; k++) {
                    const int32_t g = c->dofnode[enodes[eptr[e] + k]];
                    dst[k] = one ? g : g2l[g];
                }
                if (nen == 3) tri_gid[a++] = e;
                else quad_gid[b++] = e;
            }
        });
    }

    tm.lap("local connectivity");
    // node ids of the owned dof-nodes and the node-id span they cover (load staging)
    std::vector<int32_t> node_of_own(c->n_own);
    int64_t lo = n_nodes, hi = -1;
    for (int64_t p = 0; p < c->n_own; p++) {   // cheap: one pass over the owned nodes
        int32_t node = c->node_of_dof[c->own_begin + p];
        node_of_own[p] = node;
        lo = std::min<int64_t>(lo, node);
        hi = std::max<int64_t>(hi, node);
    }
    c->span_lo = lo;
    c->span_n = hi - lo + 1;
    FS_CUDA(c, c->d_node_of_own.alloc(c->n_own));
    FS_CUDA(c, cudaMemcpy(c->d_node_of_own.p, node_of_own.data(), sizeof(int32_t) * c->n_own, cudaMemcpyHostToDevice));
    FS_CUDA(c, c->d_stage.alloc(6 * c->span_n));
    FS_CUDA(c, c->d_F.alloc(6 * c->n_own));
    FS_CUDA(c, cudaMemset(c->d_F.p, 0, sizeof(double) * 6 * c->n_own));
    {   // the six CG vectors share one allocation (cudaMalloc is the expensive part of this step)
        const size_t len = ((size_t)6 * c->n_local + 31) & ~(size_t)31;   // 256-byte aligned pieces
        for (DevBuf<double> *v : {&c->d_b, &c->d_x, &c->d_r, &c->d_p, &c->d_q, &c->d_z}) v->release();
        FS_CUDA(c, c->d_vecpool.alloc(6 * len));
        FS_CUDA(c, cudaMemsetAsync(c->d_vecpool.p, 0, sizeof(double) * 6 * len, c->stream));
        int k = 0;
        for (DevBuf<double> *v : {&c->d_b, &c->d_x, &c->d_r, &c->d_p, &c->d_q, &c->d_z}) v->view(c->d_vecpool.p + len * (k++), 6 * (size_t)c->n_local);
    }
    c->d_full.release();
    tm.lap("vectors");
    FS_TRY(peer_window_setup(c));
    tm.lap("peer window");
    const int rcb = build_pattern(c, tri, quad, tri_gid, quad_gid);
    tm.lap("build_pattern (device)");
    return rcb;
}

// ---------------------------------------------------------------------------------------------
// loads
// ---------------------------------------------------------------------------------------------
__global__ void k_loads_from_stage(int64_t n_own, const int32_t *__restrict__ node_of_own, int64_t span_lo,
                                   const double *__restrict__ stage, double *__restrict__ F)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= 6 * n_own) return;
    int64_t p = i / 6;
    int v = (int)(i - 6 * p);
    F[i] = stage[6 * ((int64_t)node_of_own[p] - span_lo) + v];
}

int fs_set_nodal_loads(fs_context *c, const double *F)
{
    FS_CHECK_CTX(c);
    if (!c->pattern_ready) return fail(c, FS_ERR_STATE, "fs_set_nodal_loads before fs_set_mesh");
    if (!F) return fail(c, FS_ERR_ARG, "null loads");
    FS_CUDA(c, cudaSetDevice(c->device));
    FS_CUDA(c, cudaMemcpyAsync(c->d_stage.p, F + 6 * c->span_lo, sizeof(double) * 6 * c->span_n, cudaMemcpyHostToDevice, c->stream));
    k_loads_from_stage<<<nblk(6 * c->n_own, 256), 256, 0, c->stream>>>(c->n_own, c->d_node_of_own.p, c->span_lo, c->d_stage.p, c->d_F.p);
    FS_CUDA(c, cudaGetLastError());
    c->loads_set = true;
    return build_rhs(c, 1.0);
}

int fs_set_interface_loads(fs_context *c, int64_t n, const int32_t *node_ids, int dims, char dead_axis, const double *f)
{
    FS_CHECK_CTX(c);
    if (!c->pattern_ready) return fail(c, FS_ERR_STATE, "fs_set_interface_loads before fs_set_mesh");
    if (dims != 2 && dims != 3) return fail(c, FS_ERR_ARG, "dims must be 2 or 3");
    if (dims == 2 && dead_axis != 'x' && dead_axis != 'y' && dead_axis != 'z')
        return fail(c, FS_ERR_ARG, "2-D coupling needs dead axis x, y or z (fsp.cpp:92-98)");
    if (n > 0 && (!node_ids || !f)) return fail(c, FS_ERR_ARG, "null interface arrays");
    // fsp.cpp:1400-1432: only interface nodes carry load; map the 2 or 3 values onto u,v,w
    int c0 = 0, c1 = 1;
    if (dims == 2) {
        if (dead_axis == 'y') { c0 = 0; c1 = 2; }
        else if (dead_axis == 'x') { c0 = 1; c1 = 2; }
    }
    std::vector<double> stage((size_t)6 * c->span_n, 0.0);
    for (int64_t i = 0; i < n; i++) {
        int64_t node = node_ids[i];
        if (node < 0 || node >= c->n_nodes) return fail(c, FS_ERR_ARG, "interface node id out of range");
        if (node < c->span_lo || node >= c->span_lo + c->span_n) continue;
        double *dst = &stage[6 * (node - c->span_lo)];
        if (dims == 3) {
            dst[0] = f[3 * i]; dst[1] = f[3 * i + 1]; dst[2] = f[3 * i + 2];
        } else {
            dst[c0] = f[2 * i]; dst[c1] = f[2 * i + 1];
        }
    }
    FS_CUDA(c, cudaSetDevice(c->device));
    FS_CUDA(c, cudaMemcpyAsync(c->d_stage.p, stage.data(), sizeof(double) * stage.size(), cudaMemcpyHostToDevice, c->stream));
    k_loads_from_stage<<<nblk(6 * c->n_own, 256), 256, 0, c->stream>>>(c->n_own, c->d_node_of_own.p, c->span_lo, c->d_stage.p, c->d_F.p);
    FS_CUDA(c, cudaStreamSynchronize(c->stream));
    c->loads_set = true;
    return build_rhs(c, 1.0);
}

int fs_build_rhs(fs_context *c, double scale)
{
    FS_CHECK_CTX(c);
    FS_CUDA(c, cudaSetDevice(c->device));
    return build_rhs(c, scale);
}

// ---------------------------------------------------------------------------------------------
// hot path
// ---------------------------------------------------------------------------------------------
int fs_assemble(fs_context *c, float *ms)
{
    FS_CHECK_CTX(c);
    if (!c->pattern_ready) return fail(c, FS_ERR_STATE, "fs_assemble before fs_set_mesh");
    if (!c->material_set) return fail(c, FS_ERR_STATE, "fs_assemble before fs_set_material");
    FS_CUDA(c, cudaSetDevice(c->device));
    return assemble_values(c, ms);
}

int fs_get_assembly_path(fs_context *c, int *path)
{
    FS_CHECK_CTX(c);
    if (!path || !c->assembled) return fail(c, FS_ERR_STATE, "not assembled / null output");
    *path = !c->parity_valid ? 2 : ((c->asm_mode == FS_ASM_GATHER && c->gather_ready) ? 1 : 0);
    return FS_OK;
}

static void default_opts(fs_solve_opts &o)
{
    o.rtol = 1e-12;   // libMesh default TOLERANCE^2 (fs.cpp:130-133 leave it untouched)
    o.max_its = 5000;
    o.pc = FS_PC_JACOBI;
    o.norm_type = FS_NORM_PRECONDITIONED;  // KSPCG's default norm, i.e. what the reference's solve tests
    o.warm_start = 1;
    o.check_every = 0;
}

int fs_solve(fs_context *c, const fs_solve_opts *opts, fs_solve_info *info)
{
    FS_CHECK_CTX(c);
    fs_solve_opts o;
    if (opts) o = *opts;
    else default_opts(o);
    if (!(o.rtol > 0.0) || o.max_its < 0) return fail(c, FS_ERR_ARG, "bad tolerance / iteration limit");
    FS_CUDA(c, cudaSetDevice(c->device));
    return solver_run(c, &o, info);
}

__global__ void k_solution_to_nodes(int64_t n_own, const int32_t *__restrict__ node_of_own,
                                    const double *__restrict__ x_own, double *__restrict__ full)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= 6 * n_own) return;
    int64_t p = i / 6;
    int v = (int)(i - 6 * p);
    full[6 * (int64_t)node_of_own[p] + v] = x_own[i];
}

int fs_get_solution(fs_context *c, double *sols)
{
    FS_CHECK_CTX(c);
    if (!c->have_solution) return fail(c, FS_ERR_STATE, "no solution yet");
    if (!sols) return fail(c, FS_ERR_ARG, "null output");
    FS_CUDA(c, cudaSetDevice(c->device));
    if (c->d_full.n < (size_t)6 * c->n_nodes) FS_CUDA(c, c->d_full.alloc(6 * c->n_nodes));
    FS_CUDA(c, cudaMemsetAsync(c->d_full.p, 0, sizeof(double) * 6 * c->n_nodes, c->stream));
    k_solution_to_nodes<<<nblk(6 * c->n_own, 256), 256, 0, c->stream>>>(c->n_own, c->d_node_of_own.p,
                                                                         c->d_x.p + 6 * c->own_lo, c->d_full.p);
    if (c->world > 1) {
        ncclResult_t r = nccl().AllReduce(c->d_full.p, c->d_full.p, 6 * c->n_nodes, ncclDouble, ncclSum, (ncclComm_t)c->comm, c->stream);
        if (r != ncclSuccess) return fail(c, FS_ERR_COMM, std::string("ncclAllReduce: ") + nccl().GetErrorString(r));
    }
    FS_CUDA(c, cudaMemcpyAsync(sols, c->d_full.p, sizeof(double) * 6 * c->n_nodes, cudaMemcpyDeviceToHost, c->stream));
    FS_CUDA(c, cudaStreamSynchronize(c->stream));
    return FS_OK;
}

int fs_get_solution_owned(fs_context *c, int64_t *n_rows, int32_t *node_ids, double *vals)
{
    FS_CHECK_CTX(c);
    if (!c->have_solution) return fail(c, FS_ERR_STATE, "no solution yet");
    if (n_rows) *n_rows = c->n_own;
    if (node_ids)
        for (int64_t k = 0; k < c->n_own; k++) node_ids[k] = c->node_of_dof[c->own_begin + k];
    if (vals) {  // the owned rows are contiguous in the local vector layout: one copy, no kernel, no communication
        FS_CUDA(c, cudaSetDevice(c->device));
        FS_CUDA(c, cudaMemcpyAsync(vals, c->d_x.p + 6 * c->own_lo, sizeof(double) * 6 * c->n_own, cudaMemcpyDeviceToHost, c->stream));
        FS_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    return FS_OK;
}

int fs_recover_resultants(fs_context *c, double *out)
{
    FS_CHECK_CTX(c);
    if (!c->have_solution) return fail(c, FS_ERR_STATE, "no solution yet");
    if (!out) return fail(c, FS_ERR_ARG, "null output");
    FS_CUDA(c, cudaSetDevice(c->device));
    int rc = halo_exchange(c, c->d_x.p);  // the elements of the cut rows read displacements of halo nodes
    if (rc) return rc;
    DevBuf<double> res;
    FS_CUDA(c, res.alloc((size_t)6 * c->n_elem));
    FS_CUDA(c, cudaMemsetAsync(res.p, 0, sizeof(double) * 6 * c->n_elem, c->stream));
    rc = recover_resultants(c, c->d_x.p, res.p);
    if (rc) return rc;
    if (c->world > 1) {  // every element was written by exactly one rank (owner of its first node)
        ncclResult_t r = nccl().AllReduce(res.p, res.p, 6 * c->n_elem, ncclDouble, ncclSum, (ncclComm_t)c->comm, c->stream);
        if (r != ncclSuccess) return fail(c, FS_ERR_COMM, std::string("ncclAllReduce: ") + nccl().GetErrorString(r));
    }
    FS_CUDA(c, cudaMemcpyAsync(out, res.p, sizeof(double) * 6 * c->n_elem, cudaMemcpyDeviceToHost, c->stream));
    FS_CUDA(c, cudaStreamSynchronize(c->stream));
    return FS_OK;
}

int fs_solve_host(fs_context *c, const double *F, int reassemble, const fs_solve_opts *opts, double *sols,
                  fs_solve_info *info)
{
    FS_CHECK_CTX(c);
    if (F && (reassemble || !c->assembled) && c->pattern_ready) {
        // the loads travel host -> device on the copy stream WHILE the values pass runs (the two are independent:
        // the element loop needs no loads, fs.cpp:1211-1221 vs :1222-1226); the rhs is formed once both are done
        FS_CUDA(c, cudaSetDevice(c->device));
        FS_CUDA(c, cudaEventRecord(c->ev_copy, c->stream));  // earlier work on the context stream may still read the staging buffer
        FS_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->ev_copy, 0));
        FS_CUDA(c, cudaMemcpyAsync(c->d_stage.p, F + 6 * c->span_lo, sizeof(double) * 6 * c->span_n, cudaMemcpyHostToDevice, c->copy_stream));
        FS_CUDA(c, cudaEventRecord(c->ev_copy, c->copy_stream));
        int rca = fs_assemble(c, nullptr);
        if (rca) {
            cudaStreamSynchronize(c->copy_stream);  // F is borrowed for the duration of the call only
            return rca;
        }
        FS_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_copy, 0));
        k_loads_from_stage<<<nblk(6 * c->n_own, 256), 256, 0, c->stream>>>(c->n_own, c->d_node_of_own.p, c->span_lo, c->d_stage.p, c->d_F.p);
        FS_CUDA(c, cudaGetLastError());
        c->loads_set = true;
        FS_TRY(build_rhs(c, 1.0));
    } else {
        if (F) FS_TRY(fs_set_nodal_loads(c, F));
        if (reassemble || !c->assembled) FS_TRY(fs_assemble(c, nullptr));
    }
    int rc = fs_solve(c, opts, info);
    if (rc != FS_OK && rc != FS_ERR_NOT_CONVERGED) return rc;
    if (!sols) return rc;  // the caller fetches its own rows with fs_get_solution_owned
    int rc2 = fs_get_solution(c, sols);
    return rc2 != FS_OK ? rc2 : rc;
}

// ---------------------------------------------------------------------------------------------
// coupled step (fsp.cpp:257-374)
// ---------------------------------------------------------------------------------------------
int fs_interface_nodes(fs_context *c, int64_t *n, int32_t *ids)
{
    FS_CHECK_CTX(c);
    if (!c->n_nodes) return fail(c, FS_ERR_STATE, "no mesh");
    if (n) *n = (int64_t)c->iface_nodes.size();
    if (ids) memcpy(ids, c->iface_nodes.data(), sizeof(int32_t) * c->iface_nodes.size());
    return FS_OK;
}

static void axis_components(int dims, char dead_axis, int comp[3])
{
    comp[0] = 0; comp[1] = 1; comp[2] = 2;
    if (dims == 2) {
        if (dead_axis == 'y') { comp[0] = 0; comp[1] = 2; }
        else if (dead_axis == 'x') { comp[0] = 1; comp[1] = 2; }
    }
}

// sols / preSols of the coupling loop (fsp.cpp:77-87), zero until the first step like the reference's new[]-and-fill
static void coupled_state(fs_context *c)
{
    const size_t n = (size_t)6 * c->n_nodes;
    if (c->sols.size() != n) c->sols.assign(n, 0.0);
    if (c->pre_sols.size() != n) c->pre_sols.assign(n, 0.0);
}

int fs_step(fs_context *c, int dims, char dead_axis, const double *forces_in, const fs_solve_opts *opts,
            double *displ_out, fs_solve_info *info)
{
    FS_CHECK_CTX(c);
    if (!forces_in || !displ_out) return fail(c, FS_ERR_ARG, "null coupling arrays");
    const int64_t nif = (int64_t)c->iface_nodes.size();
    FS_TRY(fs_set_interface_loads(c, nif, c->iface_nodes.data(), dims, dead_axis, forces_in));
    if (!c->assembled) FS_TRY(fs_assemble(c, nullptr));  // K is constant across coupling iterations
    int rc = fs_solve(c, opts, info);
    if (rc != FS_OK && rc != FS_ERR_NOT_CONVERGED) return rc;
    coupled_state(c);
    FS_TRY(fs_get_solution(c, c->sols.data()));
    int comp[3];
    axis_components(dims, dead_axis, comp);
    for (int64_t i = 0; i < nif; i++) {  // fsp.cpp:286-317
        const int64_t id = c->iface_nodes[i];
        for (int d = 0; d < dims; d++) displ_out[i * dims + d] = c->sols[6 * id + comp[d]] - c->pre_sols[6 * id + comp[d]];
    }
    return rc;
}

int fs_commit_step(fs_context *c, int dims, char dead_axis)
{
    FS_CHECK_CTX(c);
    if (dims != 2 && dims != 3) return fail(c, FS_ERR_ARG, "dims must be 2 or 3");
    int comp[3];
    axis_components(dims, dead_axis, comp);
    coupled_state(c);
    for (int32_t id : c->iface_nodes)  // fsp.cpp:347-368
        for (int d = 0; d < dims; d++) c->pre_sols[6 * (int64_t)id + comp[d]] = c->sols[6 * (int64_t)id + comp[d]];
    return FS_OK;
}

// ---------------------------------------------------------------------------------------------
// parity / inspection
// ---------------------------------------------------------------------------------------------
int fs_get_sizes(fs_context *c, int64_t *n_dofnodes, int64_t *n_blocks, int64_t *n_colors, int64_t *own_begin, int64_t *own_end)
{
    FS_CHECK_CTX(c);
    if (!c->pattern_ready) return fail(c, FS_ERR_STATE, "no mesh");
    if (n_dofnodes) *n_dofnodes = c->n_dofnodes_global;
    if (n_blocks) *n_blocks = c->n_blocks;
    if (n_colors) *n_colors = c->n_colors;
    if (own_begin) *own_begin = c->own_begin;
    if (own_end) *own_end = c->own_end;
    return FS_OK;
}

int fs_export_dof_order(fs_context *c, int32_t *dofnode)
{
    FS_CHECK_CTX(c);
    if (!c->n_nodes || !dofnode) return fail(c, FS_ERR_STATE, "no mesh");
    memcpy(dofnode, c->dofnode.data(), sizeof(int32_t) * c->n_nodes);
    return FS_OK;
}

int fs_export_csr(fs_context *c, int64_t *rowptr, int32_t *colidx, double *vals)
{
    FS_CHECK_CTX(c);
    if (!c->pattern_ready) return fail(c, FS_ERR_STATE, "no pattern");
    if (vals && !c->assembled) return fail(c, FS_ERR_STATE, "values requested before fs_assemble");
    FS_CUDA(c, cudaSetDevice(c->device));
    std::vector<int32_t> nptr(c->n_own + 1), nadj(c->n_blocks);
    FS_CUDA(c, cudaMemcpy(nptr.data(), c->d_nptr.p, sizeof(int32_t) * (c->n_own + 1), cudaMemcpyDeviceToHost));
    FS_CUDA(c, cudaMemcpy(nadj.data(), c->d_nadj.p, sizeof(int32_t) * c->n_blocks, cudaMemcpyDeviceToHost));
    for (int64_t p = 0; p < c->n_own; p++) {
        const int64_t deg = nptr[p + 1] - nptr[p];
        for (int a = 0; a < 6; a++) {
            const int64_t base = 36 * (int64_t)nptr[p] + a * 6 * deg;
            if (rowptr) rowptr[6 * p + a] = base;
            if (colidx)
                for (int64_t j = 0; j < deg; j++)
                    for (int b = 0; b < 6; b++)
                        colidx[base + 6 * j + b] = 6 * c->local_to_global[nadj[nptr[p] + j]] + b;
        }
    }
    if (rowptr) rowptr[6 * c->n_own] = 36 * (int64_t)c->n_blocks;
    if (vals) {
        FS_TRY(ensure_parity_values(c));  // shells in the xy plane are assembled into the compacted SpMV format only
        FS_CUDA(c, cudaMemcpy(vals, c->d_vals.p, sizeof(double) * 36 * c->n_blocks, cudaMemcpyDeviceToHost));
    }
    return FS_OK;
}

int fs_export_rhs(fs_context *c, double *rhs)
{
    FS_CHECK_CTX(c);
    if (!c->rhs_ready) return fail(c, FS_ERR_STATE, "no rhs");
    FS_CUDA(c, cudaSetDevice(c->device));
    FS_CUDA(c, cudaStreamSynchronize(c->stream));
    FS_CUDA(c, cudaMemcpy(rhs, c->d_b.p + 6 * c->own_lo, sizeof(double) * 6 * c->n_own, cudaMemcpyDeviceToHost));
    return FS_OK;
}

int fs_debug_element_matrices(fs_context *c, double *out)
{
    FS_CHECK_CTX(c);
    if (!c->pattern_ready || !c->material_set) return fail(c, FS_ERR_STATE, "mesh and material required");
    if (c->world != 1) return fail(c, FS_ERR_STATE, "single-rank only");
    FS_CUDA(c, cudaSetDevice(c->device));
    return debug_element_matrices(c, out);
}

int fs_spmv_host(fs_context *c, const double *x, double *y)
{
    FS_CHECK_CTX(c);
    if (!c->assembled) return fail(c, FS_ERR_STATE, "not assembled");
    if (c->world != 1) return fail(c, FS_ERR_STATE, "single-rank only");
    FS_CUDA(c, cudaSetDevice(c->device));
    FS_CUDA(c, cudaMemcpyAsync(c->d_p.p, x, sizeof(double) * 6 * c->n_own, cudaMemcpyHostToDevice, c->stream));
    FS_TRY(spmv_once(c, c->d_p.p, c->d_q.p));
    FS_CUDA(c, cudaMemcpyAsync(y, c->d_q.p, sizeof(double) * 6 * c->n_own, cudaMemcpyDeviceToHost, c->stream));
    FS_CUDA(c, cudaStreamSynchronize(c->stream));
    return FS_OK;
}

int fs_bench_spmv(fs_context *c, int reps, fs_solve_info *info)
{
    FS_CHECK_CTX(c);
    if (!c->assembled || reps <= 0 || !info) return fail(c, FS_ERR_STATE, "not assembled / bad reps");
    FS_CUDA(c, cudaSetDevice(c->device));
    return spmv_kernel_time(c, reps, &info->spmv_ms, &info->solve_ms);   // solve_ms: the same kernel reading p (peer window)
}

int fs_set_ml_options(fs_context *c, int64_t max_points, int dense_points, int gamma)
{
    FS_CHECK_CTX(c);
    // gamma: decimal digits, most significant first: digit l = visits of lattice level l+1 per visit of level l; the last
    // digit also serves every deeper level (2 = W everywhere, 21 = W on top and V below, 2211 = W on the two finest lattices)
    bool digits_ok = gamma >= 1 && gamma <= 333333;
    for (int g = gamma; g > 0 && digits_ok; g /= 10) digits_ok = g % 10 >= 1 && g % 10 <= 3;
    if (max_points < 1 || dense_points < 1 || dense_points > fs::ML_DENSE_MAX_POINTS || !digits_ok)
        return fail(c, FS_ERR_ARG, "multilevel options out of range");
    if (max_points != c->ml_max_points || dense_points != c->ml_dense_points || gamma != c->ml_gamma) {
        c->ml_max_points = max_points;
        c->ml_dense_points = dense_points;
        c->ml_gamma = gamma;
        c->ml_geom_ready = c->ml_values_ready = false;
        if (c->cg_graph_exec) { cudaGraphExecDestroy(c->cg_graph_exec); c->cg_graph_exec = nullptr; }
    }
    return FS_OK;
}

int fs_get_ml_info(fs_context *c, int64_t *levels, int64_t *cells, double *weights, double *setup_ms)
{
    FS_CHECK_CTX(c);
    if (!levels) return fail(c, FS_ERR_ARG, "null levels");
    *levels = c->ml_geom_ready ? c->ml.n_lat : 0;
    for (int l = 0; l < (c->ml_geom_ready ? c->ml.n_lat : 0); l++) {
        if (cells)
            for (int d = 0; d < 3; d++) cells[3 * l + d] = c->ml.lat[l].g.np[d];
        if (weights) weights[l + 1] = c->ml.lat[l].lambda;
    }
    if (weights) weights[0] = c->ml.lambda0;
    if (setup_ms) *setup_ms = c->ml_values_ready ? c->ml.setup_ms : 0.0;
    return FS_OK;
}

int fs_get_ml_dist_levels(fs_context *c, int64_t *n_dist)
{
    FS_CHECK_CTX(c);
    if (!n_dist) return fail(c, FS_ERR_ARG, "null output");
    *n_dist = c->ml_geom_ready ? c->ml.n_dist : 0;
    return FS_OK;
}

int fs_get_ml_compact_levels(fs_context *c, int64_t *n_compact)
{
    FS_CHECK_CTX(c);
    if (!n_compact) return fail(c, FS_ERR_ARG, "null output");
    *n_compact = 0;
    if (c->ml_values_ready)
        for (int l = 0; l < c->ml.n_lat; l++) *n_compact += c->ml.lat[l].compact_kind >= 0 ? 1 : 0;
    return FS_OK;
}

int fs_get_ml_profile(fs_context *c, double ms[8], int64_t *iterations, int reset)
{
    FS_CHECK_CTX(c);
    if (!ms || !iterations) return fail(c, FS_ERR_ARG, "null output");
    for (int k = 0; k < 8; k++) ms[k] = c->prof.ms[k];
    *iterations = c->prof.n;
    if (reset) {
        for (double &v : c->prof.ms) v = 0.0;
        c->prof.n = 0;
    }
    return FS_OK;
}

int fs_debug_ml_level(fs_context *c, int level, int what, double *out, int64_t capacity, int64_t *count)
{
    FS_CHECK_CTX(c);
    if (!c->ml_values_ready) return fail(c, FS_ERR_STATE, "multilevel preconditioner not set up");
    if (level < 0 || level >= c->ml.n_lat || !count) return fail(c, FS_ERR_ARG, "bad level");
    fs::MlLevelBuf &L = c->ml.lat[level];
    const int64_t n6 = 6 * (int64_t)L.g.n;
    const double *src = nullptr;
    int64_t n = 0;
    if (what == 0) { src = L.A.p; n = (int64_t)L.g.ns * 6 * n6; }
    else if (what == 1) { src = L.dinv.p; n = 6 * n6; }
    else if (what == 2 && L.dense) { src = L.minv.p; n = n6 * n6; }
    else return fail(c, FS_ERR_ARG, "nothing of that kind on this level");
    *count = n;
    if (!out) return FS_OK;
    if (capacity < n) return fail(c, FS_ERR_ARG, "buffer too small");
    FS_CUDA(c, cudaSetDevice(c->device));
    FS_CUDA(c, cudaMemcpy(out, src, sizeof(double) * n, cudaMemcpyDeviceToHost));
    return FS_OK;
}

int fs_apply_mlrbm_host(fs_context *c, const double *r, double *z)
{
    FS_CHECK_CTX(c);
    if (!c->assembled) return fail(c, FS_ERR_STATE, "not assembled");
    if (c->world != 1) return fail(c, FS_ERR_STATE, "single-rank only");
    if (!r || !z) return fail(c, FS_ERR_ARG, "null vector");
    FS_CUDA(c, cudaSetDevice(c->device));
    FS_CUDA(c, cudaMemcpyAsync(c->d_r.p, r, sizeof(double) * 6 * c->n_own, cudaMemcpyHostToDevice, c->stream));
    FS_TRY(pc_apply_mlrbm_once(c));
    FS_CUDA(c, cudaMemcpyAsync(z, c->d_z.p, sizeof(double) * 6 * c->n_own, cudaMemcpyDeviceToHost, c->stream));
    FS_CUDA(c, cudaStreamSynchronize(c->stream));
    return FS_OK;
}

}  // extern "C"

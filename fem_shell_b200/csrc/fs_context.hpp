// fs_context.hpp -- internal state behind the opaque fs_context of include/femshell_b200.h
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <string>
#include <vector>

#include "../../include/femshell_b200.h"
#include "fs_gather_plan.hpp"
#include "fs_mlpc.cuh"

struct ncclComm;
namespace fs { struct PeerWin; }

namespace fs {

template <class T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    bool owned = true;
    ~DevBuf() { release(); }
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    void release()
    {
        if (p && owned) cudaFree(p);
        p = nullptr;
        n = 0;
        owned = true;
    }
    void view(T *ptr, size_t count)  // non-owning window into another allocation
    {
        release();
        p = ptr;
        n = count;
        owned = false;
    }
    cudaError_t alloc(size_t count)
    {
        release();
        if (count == 0) count = 1;
        cudaError_t e = cudaMalloc((void **)&p, count * sizeof(T));
        if (e == cudaSuccess) n = count;
        return e;
    }
};

// CG scalars living on the device; one cache line, read by every kernel of an iteration
struct CgState {
    double rz;        // r.z of the current iterate
    double pq;        // p.Ap
    double alpha;     // rz / pq
    double beta;      // rz_new / rz
    double nrm2;      // ||r||^2 or ||z||^2 after the update (norm_type)
    double bnorm2;    // ||b||^2 or ||M^-1 b||^2
    double tol2;      // (rtol)^2
    long long iter;
    long long max_its;
    int done;         // 1 -> all later kernels of the batch are no-ops
    int status;       // FS_OK / FS_ERR_NOT_CONVERGED / FS_ERR_BREAKDOWN
};

struct Peer {
    int rank = -1;
    int64_t send_count = 0;   // nodes
    int64_t recv_count = 0;   // nodes
    int64_t send_off = 0;     // offset (nodes) into the packed send buffer / send index list
    int64_t recv_off = 0;     // offset (nodes) into the halo segment of the local vector
};

// one lattice level of the multilevel preconditioner (fs_mlpc.cuh)
struct MlLevelBuf {
    LatGeom g;
    bool dense = false;          // coarsest level: minv holds the dense pseudo-inverse
    DevBuf<double> A;            // ns * 36 * n stencil values (structure of arrays, fs_mlpc.cu)
    DevBuf<double> dinv;         // 36 * n pseudo-inverses of the diagonal blocks
    DevBuf<double> minv;         // (6n)^2, dense level only
    // shells in a coordinate plane: in-plane and out-of-plane modes never couple, on any level.  Ac / Dc hold the 18
    // structurally non-zero entries of every 6x6 block, one plane of n values per (slot, item): thread = cell streams
    // them with unit stride (k_lat_stencil_c); A / dinv stay as the probing wrote them (fs_debug_ml_level)
    DevBuf<double> Ac, Dc;       // ns * 18 * n, 18 * n
    int compact_kind = -1;       // -1: not compacted; else the mask kind (0 xy, 1 xz, 2 yz)
    DevBuf<double> x, xb, b, r, t;  // 6n each
    double omega = 0.0, lambda = 0.0;
    // distribution over the ranks (fs_mlpc.cu "distributed lattice levels"): cells are dealt out in slabs (all cells
    // sharing the index of the slowest-varying active dimension); arrays keep their full size, a rank computes its own
    // slabs [s0, s1) and keeps `halo` slabs of the neighbours' values valid on either side
    bool dist = false;
    int slab_len = 0, n_slabs = 0, s0 = 0, s1 = 0, halo = 1;
    int c0 = 0, c1 = 0;             // owned cells [c0, c1) (replicated level: all)
    DevBuf<double> stage;           // one slab: neighbour's partial sums of a restriction
};

struct MlHier {
    int n_lat = 0;
    MlLevelBuf lat[ML_MAX_LEVELS];
    DevBuf<int32_t> d_agg;       // n_local: cell of the first lattice holding each local node
    DevBuf<int32_t> d_sup_ptr, d_sup_node;  // cell -> owned nodes (owned-relative ids)
    DevBuf<double> d_r1, d_t;    // 6 * n_local work vectors of the mesh level
    // shells in a coordinate plane: the inverses of the 6x6 diagonal blocks (d_minv, 36 per node) have the 14-entry pattern
    // of the mesh blocks; the four mesh-level kernels of the cycle that apply them read this copy (14 per node, row-major
    // over the set bits of the mask), bit-identical sums.  dinv_mask = 0: not compacted
    DevBuf<double> d_dinv_c;
    unsigned long long dinv_mask = 0;
    DevBuf<double> d_scalar;
    double omega0 = 0.0, lambda0 = 0.0;
    float setup_ms = 0.f;
    int n_dist = 0;              // leading lattice levels that are distributed (0: all replicated)
    std::vector<int> bounds[ML_MAX_LEVELS];  // first owned slab of every rank (+ n_slabs), per distributed level
};

}  // namespace fs

struct fs_context {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaStream_t copy_stream = nullptr;    // host -> device load copies that overlap the values pass (fs_solve_host)
    cudaEvent_t ev_copy = nullptr;
    std::string err;

    // distributed
    int rank = 0, world = 1;
    ncclComm *comm = nullptr;

    // NVLink peer window (fs_peer.cuh): [mailbox | p] mapped into every rank of the box
    int comm_pref = FS_COMM_AUTO;
    bool peer_ready = false;
    void *win_base = nullptr;                  // this rank's window (cudaMalloc)
    void *peer_base[8] = {};                   // the other ranks' windows (cudaIpcOpenMemHandle)
    fs::DevBuf<fs::PeerWin> d_pw;
    fs::DevBuf<int32_t> d_push_peer, d_push_dst;
    fs::DevBuf<uint8_t> d_is_send;             // n_own: node is on the send list (folded halo push, k_direction)
    bool push_foldable = false;                // every send-list node goes to exactly one neighbour

    // material / switches
    double nu = 0.3, E = 1e7, thickness = 1.0;
    bool material_set = false;
    int quirks = FS_QUIRKS_REFERENCE;
    int dof_mode = FS_DOF_FIRST_ENCOUNTER;
    int asm_mode = FS_ASM_GATHER;

    // host copy of the (replicated) mesh description that later calls need
    int64_t n_nodes = 0, n_elem = 0;
    std::vector<int32_t> dofnode;        // node id -> global dof-node (-1 orphan)
    std::vector<int32_t> node_of_dof;    // global dof-node -> node id
    std::vector<uint8_t> node_mask;      // Dirichlet bits per node id
    std::vector<int32_t> iface_nodes;    // coupling interface nodes (ids 2,20,21), ascending
    int64_t n_dofnodes_global = 0;

    // rank-local numbering: local nodes = owned + halo dof-nodes sorted by GLOBAL dof-node id, so the
    // owned range [own_begin, own_end) is the contiguous local range [own_lo, own_lo + n_own) and
    // local order == global order (the CSR columns stay sorted without a second key)
    int64_t own_begin = 0, own_end = 0, n_own = 0, n_local = 0, own_lo = 0;
    std::vector<int32_t> local_to_global;  // local dof-node -> global dof-node
    int64_t span_lo = 0, span_n = 0;       // node-id span covering the owned nodes (load staging)
    std::vector<fs::Peer> peers;
    int64_t send_total = 0;

    // device mesh (local numbering)
    int64_t n_tri = 0, n_quad = 0;         // local elements (touching owned nodes)
    fs::DevBuf<double> d_xyz;              // n_local * 3
    fs::DevBuf<uint8_t> d_mask;            // n_local
    fs::DevBuf<int32_t> d_tri, d_quad;     // connectivity, local dof-node ids, sorted by colour
    fs::DevBuf<int32_t> d_tri_pos, d_quad_pos;  // block slot of node j in the row of node i (or -1)
    fs::DevBuf<int32_t> d_tri_gid, d_quad_gid;  // original element id (debug / determinism)
    std::vector<int64_t> tri_color_off, quad_color_off;  // n_colors+1 offsets
    int64_t n_colors = 0;
    bool colored = false;                  // colouring + colour-sorted element arrays exist (built by the first coloured pass)
    // row-gather assembly schedule (fs_assembly.cu, build_gather_schedule)
    fs::DevBuf<fs::GatherChunk> d_g_chunks;
    fs::DevBuf<int4> d_g_info, d_g_nodes;  // packed thread table: {meta, row info, Dirichlet bits, slots}, node ids
    fs::DevBuf<double> d_qgp;              // 96 doubles: Gauss-point shape-derivative table (fs_elements.cuh QuadGpTab)
    int64_t n_g_chunks = 0;
    bool gather_ready = false, gather_unavailable = false;
    // slice pass (fs_slice_asm.cu): shells in the xy plane are assembled straight into the sliced-ELL SpMV format
    fs::DevBuf<int32_t> d_sl_ptr;          // n_own+1: first incidence record of every owned row (slice s starts at row 32 s)
    fs::DevBuf<int4> d_sl_info, d_sl_nodes;  // packed thread table, one record per (element, node row) incidence
    fs::DevBuf<int2> d_sl_meta;            // per slice: {emit phases, element kinds present}
    // planar shells (any orientation): Q = rows x^, y^, n of a fixed frame whose third axis is the plane normal.  In that
    // frame the shell lies in "its xy plane": the slice pass assembles Q~ K Q~^T (Q~ = diag(Q, Q) per node, 14 of 36
    // entries per block) from node coordinates rotated into the frame, and the SpMV applies Q~ / Q~^T on the fly.
    bool planar = false, plane_rot = false;   // plane_rot: Q is not the identity
    double plane_Q[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    fs::DevBuf<double> d_xyz_plane;        // n_local * 3: Q x, third component snapped to the plane's constant (plane_rot only)
    bool slice_ready = false;              // the plan exists for this mesh (planar, rows fit)
    size_t slice_smem = 0;
    int slice_threads = 128;
    bool parity_valid = false;             // d_vals holds the values of the current assembly (else: formed on demand)

    // block-CSR matrix of the owned rows: row 6p+a occupies vals[36*nptr[p] + a*6*deg ...]
    fs::DevBuf<int32_t> d_nptr;            // n_own+1
    fs::DevBuf<int32_t> d_nadj;            // n_blocks, LOCAL column dof-nodes, ascending in GLOBAL order
    fs::DevBuf<double> d_vals;             // 36*n_blocks
    int64_t n_blocks = 0;
    bool pattern_ready = false, assembled = false;

    // zero-compacted SpMV copy of the matrix (fs_sell.cuh); rebuilt lazily after every values pass
    int spmv_format_pref = FS_SPMV_AUTO;
    bool sell_checked = false;             // the current d_vals have been inspected
    bool sell_active = false;              // the SpMV runs on the compacted copy
    bool sell_layout_ready = false;        // slice pointers / columns built for this mesh
    unsigned long long sell_detected = 0;  // union pattern of all 6x6 blocks (bit 6a+b)
    unsigned long long sell_mask = 0;      // the kernel mask in use (superset of sell_detected)
    int sell_nz = 36, sell_kind = -1;
    int64_t sell_slices = 0, sell_slots = 0;
    int sell_dmax_max = 0;                 // widest slice (blocks per row): sizes the shared memory of k_sell_fill_t
    fs::DevBuf<int32_t> d_sell_sptr, d_sell_adj;
    fs::DevBuf<int32_t> d_sell_hflag;      // several ranks, per slice: 1 = a row of the slice reads halo blocks (fs_sell.cuh)
    fs::DevBuf<double> d_sell_vals;
    fs::DevBuf<unsigned long long> d_sell_mask;
    int sell_blocks_per_sm = 2;

    // loads / vectors (length 6*n_local unless noted)
    fs::DevBuf<double> d_F;                // 6*n_own loads in dof order (unconstrained values)
    fs::DevBuf<double> d_stage;            // 6*span_n node-ordered staging for loads / solution
    fs::DevBuf<int32_t> d_node_of_own;     // n_own: mesh node id of owned dof-node p
    fs::DevBuf<double> d_full;             // 6*n_nodes gather buffer (multi-rank solution)
    fs::DevBuf<double> d_vecpool;          // one allocation behind the six vectors below
    fs::DevBuf<double> d_b, d_x, d_r, d_p, d_q, d_z;
    fs::DevBuf<double> d_minv;             // 6*n_own (Jacobi) or 36*n_own (block)
    int minv_kind = -1;

    // multilevel rigid-body-mode preconditioner (fs_mlpc.cuh / fs_mlpc.cu)
    double bbox_lo[3] = {0, 0, 0}, bbox_hi[3] = {0, 0, 0};  // of all mesh nodes (identical on every rank)
    double ml_h[3] = {0, 0, 0};            // largest element extent per axis (identical on every rank; measured by ml_build_geometry)
    int64_t ml_max_points = 1 << 22;       // cap on the cells of the first lattice
    int ml_dense_points = fs::ML_DENSE_DEFAULT_POINTS;  // a lattice with at most this many cells is solved densely
    int ml_gamma = 2;                      // cycle index on the lattice levels (1 = V, 2 = W; two digits: first lattice, deeper ones)
    bool ml_geom_ready = false, ml_values_ready = false;
    fs::MlHier ml;
    fs::DevBuf<double> d_partials;         // per-block partial sums
    fs::DevBuf<fs::CgState> d_state;
    fs::DevBuf<unsigned int> d_counter;
    fs::DevBuf<int> d_flag;                // scratch error flag
    fs::DevBuf<double> d_sendbuf;          // 6*send_total
    fs::DevBuf<int32_t> d_send_idx;        // send_total local node ids
    fs::CgState *h_state = nullptr;        // pinned
    cudaGraphExec_t cg_graph_exec = nullptr;  // captured CG iterations (invalidated with the mesh / preconditioner)
    int cg_graph_key = -1;
    double *cg_graph_red = nullptr;
    bool loads_set = false, rhs_ready = false, have_solution = false;

    // FS_ML_PROFILE=1 (lab): the multilevel-preconditioned iterations run eagerly with events between their stages
    struct StageProf {
        bool on = false;
        cudaEvent_t ev[12] = {};
        double ms[10] = {};
        long n = 0;
    } prof;

    // coupled step
    std::vector<double> sols, pre_sols;

    int sm_count = 148;
    int spmv_blocks_per_sm = 4, vec_blocks_per_sm = 4;
};

namespace fs {

// FS_TIMING=1: phase timings of the set-up paths on stderr (host wall clock; the caller synchronises where it matters)
struct PhaseTimer {
    bool on;
    double t0;
    const char *what;
    static double now()
    {
        timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        return ts.tv_sec + 1e-9 * ts.tv_nsec;
    }
    explicit PhaseTimer(const char *w) : on(getenv("FS_TIMING") != nullptr), t0(0.0), what(w) { if (on) t0 = now(); }
    void lap(const char *phase)
    {
        if (!on) return;
        const double t = now();
        fprintf(stderr, "[fs timing] %s / %-28s %8.2f ms\n", what, phase, 1e3 * (t - t0));
        t0 = t;
    }
};

inline int fail(fs_context *c, int code, const std::string &msg)
{
    if (c) c->err = msg;
    return code;
}

#define FS_CUDA(ctx, call)                                                                     \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess)                                                                \
            return fs::fail(ctx, FS_ERR_CUDA,                                                  \
                            std::string(#call) + ": " + cudaGetErrorString(e__) + " (" +      \
                                __FILE__ + ":" + std::to_string(__LINE__) + ")");              \
    } while (0)

// assembly.cu
int upload_element_constants(fs_context *c);
int build_pattern(fs_context *c, const std::vector<int32_t> &tri, const std::vector<int32_t> &quad,
                  const std::vector<int32_t> &tri_gid, const std::vector<int32_t> &quad_gid);
int assemble_values(fs_context *c, float *ms);
int build_gather_schedule(fs_context *c);
int ensure_coloring(fs_context *c);
int ensure_parity_values(fs_context *c);   // d_vals <- the current assembly (no-op when it already holds it)
// slice_asm.cu
int sell_layout_build(fs_context *c);
int slice_plan_build(fs_context *c);
int assemble_slice_enqueue(fs_context *c);
int extract_minv_sell(fs_context *c, int pc, int *d_bad);
int build_rhs(fs_context *c, double scale);
int debug_element_matrices(fs_context *c, double *out_host);
int recover_resultants(fs_context *c, const double *d_x, double *d_out);

// solver.cu
int solver_query_occupancy(fs_context *c);
int solver_prepare(fs_context *c, int pc);
int solver_run(fs_context *c, const fs_solve_opts *o, fs_solve_info *info);
int spmv_once(fs_context *c, const double *d_in, double *d_out, bool check_done = false);   // halo exchange + SpMV
int spmv_local(fs_context *c, const double *d_in, double *d_out, bool check_done = false);  // SpMV only (halo already valid)
int halo_exchange(fs_context *c, double *d_vec);
int spmv_format_prepare(fs_context *c);
int pc_apply_mlrbm_once(fs_context *c);
int spmv_kernel_time(fs_context *c, int reps, float *ms_per_launch, float *ms_on_p = nullptr);

// mlpc.cu
int ml_prepare(fs_context *c);
int ml_enqueue_apply(fs_context *c, bool init, double *red, int fin, int vec_grid);
int ml_ensure_partials(fs_context *c);

// peer.cu
int peer_window_setup(fs_context *c);     // collective; call after the vectors are sized
int peer_window_teardown(fs_context *c);  // collective

}  // namespace fs

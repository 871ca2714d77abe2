/*
 * femshell_b200.h -- C ABI of the B200-native fem-shell hot path
 * (flat-shell element-stiffness assembly -> sparse PCG solve).
 *
 * Every entry point replaces one piece of the reference's path through
 * libMesh/PETSc (reference = precice/fem-shell; "fs.cpp" =
 * src/fem-shell/fem-shell.cpp, "fsp.cpp" =
 * src/fem-shell/preCICE/fem-shell_precice.cpp).  The reference has no FFI of
 * its own: its boundary is the libMesh assemble callback
 *     void assemble_elasticity(EquationSystems&, const std::string&)
 * (src/fem-shell/fem-shell.h:75, fs.cpp:1160) registered at fs.cpp:85 and
 * driven by equation_systems.solve() (fs.cpp:138, fsp.cpp:271).  A maintainer
 * binds this library by replacing those calls; INTEGRATION.md shows the stub.
 *
 * Conventions: plain pointers and sizes only; all host arrays are borrowed for
 * the duration of the call; every function returns 0 (FS_OK) or a negative
 * fs_status and leaves a message retrievable with fs_last_error(); no C++
 * exception crosses this boundary.  One context = one GPU = one CUDA stream;
 * calls on one context must be serialised by the caller; different contexts
 * may be driven from different host threads (the passes that use the
 * per-device element constants serialise themselves).  There is no CPU
 * fallback: without a CUDA device fs_create fails.
 */
#ifndef FEMSHELL_B200_H
#define FEMSHELL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fs_context fs_context;

typedef enum fs_status {
    FS_OK = 0,
    FS_ERR_ARG = -1,            /* bad argument / malformed mesh */
    FS_ERR_CUDA = -2,           /* CUDA runtime error (message has the detail) */
    FS_ERR_STATE = -3,          /* call order violated (e.g. solve before assemble) */
    FS_ERR_NOT_CONVERGED = -4,  /* max_its reached before rtol */
    FS_ERR_BREAKDOWN = -5,      /* p.Ap <= 0 : matrix not positive definite */
    FS_ERR_COMM = -6,           /* NCCL error */
    FS_ERR_IO = -7              /* file could not be read / written */
} fs_status;

/* XDA element type ids (src/meshgen/main_all.cpp:245-247; libMesh TRI3 / QUAD4) */
enum { FS_TRI3 = 3, FS_QUAD4 = 5 };

/* DOF numbering of the assembled system (libMesh DofMap, call site fs.cpp:125,1205).
 * FIRST_ENCOUNTER = libMesh's var-major distribution: node bases in first-encounter order
 * over elements in id order.  NODE_ID = 6*node_id+var.  SURVEY.md section 8a (a12). */
enum { FS_DOF_FIRST_ENCOUNTER = 0, FS_DOF_NODE_ID = 1 };

/* -pc_type none | jacobi | pbjacobi (PETSc flags passed through by the reference, doc/implementation.tex:68-72) */
enum { FS_PC_NONE = 0, FS_PC_JACOBI = 1, FS_PC_BJACOBI6 = 2,
       /* beyond the reference's documented options: one smoothed-aggregation multigrid cycle with rigid-body-mode
        * coarse spaces on nested lattices over the mesh (fem_shell_b200/csrc/fs_mlpc.cuh).  Same Krylov method,
        * same converged displacements, O(1000x) fewer iterations on large plates; unpreconditioned norm only. */
       FS_PC_MLRBM = 3 };

/* convergence norm: ||r||/||b|| or PETSc's default preconditioned ||M^-1 r||/||M^-1 b|| */
enum { FS_NORM_UNPRECONDITIONED = 0, FS_NORM_PRECONDITIONED = 1 };

/* arithmetic quirks of the reference, reproduced by default */
enum { FS_QUIRK_Y21 = 1 /* fs.cpp:586 */, FS_QUIRK_DET_LU = 2 /* fs.cpp:512,652 */, FS_QUIRKS_REFERENCE = 3 };

/* assembly strategy: row-gather (default: every CSR value written once, deterministic, no atomics;
 * falls back to the coloured pass if one block row exceeds a warp's shared-memory slice) or
 * graph-coloured scatter-add (one launch per colour, atomics-free read-modify-write) */
enum { FS_ASM_COLORED = 0, FS_ASM_GATHER = 1 };

/* storage the SpMV streams: AUTO = after each values pass measure the union pattern of all 6x6 blocks and,
 * when it is one of the planar-shell patterns (membrane / bending / drilling decoupled: 14 of 36 entries),
 * iterate on a zero-compacted sliced-ELL copy; FULL = always the parity format (explicit zeros included).
 * fs_export_csr returns the parity format either way. */
enum { FS_SPMV_AUTO = 0, FS_SPMV_FULL = 1 };

/* how the CG iteration exchanges data between the GPUs of one box: PEER = the kernels push halos and
 * partial dot products straight into the other ranks' memory over NVLink (cudaIpc windows, no NCCL call
 * in the loop); NCCL = ncclSend/Recv + ncclAllReduce per iteration; AUTO = PEER when every rank could map
 * every other rank's window, else NCCL.  Must be set identically on all ranks before fs_set_mesh. */
enum { FS_COMM_AUTO = 0, FS_COMM_NCCL = 1, FS_COMM_PEER = 2 };

typedef struct fs_solve_opts {
    double rtol;        /* relative tolerance (reference default 1e-12 = TOLERANCE^2, fs.cpp:130-133) */
    int64_t max_its;    /* reference default 5000 */
    int pc;             /* FS_PC_* */
    int norm_type;      /* FS_NORM_* */
    int warm_start;     /* 1: start from the previous solution (libMesh/PETSc default, fsp.cpp:271) */
    int check_every;    /* iterations enqueued between host polls of the device flag; <=0 -> default */
} fs_solve_opts;

typedef struct fs_solve_info {
    int64_t iterations;
    double rel_residual;  /* in the norm selected by norm_type */
    int status;           /* FS_OK, FS_ERR_NOT_CONVERGED or FS_ERR_BREAKDOWN */
    float solve_ms;       /* device time of the Krylov loop (CUDA events on the context stream) */
    float spmv_ms;        /* filled by fs_bench_spmv only */
} fs_solve_info;

/* ---- life cycle -------------------------------------------------------- */
/* replaces LibMeshInit (fs.cpp:28): binds the context to CUDA device `device` */
int fs_create(fs_context **ctx, int device);
int fs_destroy(fs_context *ctx);
const char *fs_last_error(const fs_context *ctx);
/* CUDA stream of the context as a cudaStream_t (for callers timing with their own events) */
void *fs_get_stream(fs_context *ctx);

/* ---- multi-GPU: one process per GPU ------------------------------------ */
/* Call BEFORE fs_set_mesh.  nccl_unique_id = the 128 bytes of an ncclUniqueId created on
 * rank 0 (fs_dist_unique_id) and broadcast by the host program.  Replaces the MPI
 * communicator libMesh/PETSc use (fs.cpp:28,35).  Collective: creates the NCCL communicator and runs one
 * all-reduce, all-gather and neighbour send/recv so that NCCL's channel set-up (0.3 s on 2 GPUs, about a
 * second on 8) is paid here and not by the first fs_set_mesh. */
int fs_dist_unique_id(uint8_t id_out[128]);
int fs_dist_init(fs_context *ctx, int rank, int world, const uint8_t nccl_unique_id[128]);
int fs_set_comm_mode(fs_context *ctx, int mode);
/* FS_COMM_NCCL or FS_COMM_PEER: what the iteration of the current mesh uses (single rank: FS_COMM_NCCL) */
int fs_get_comm_mode(fs_context *ctx, int *mode);
/* FS_COMM_PEER: what the CG kernels of this rank spent WAITING for the other GPUs since the last reset, measured
 * with clock64: out = {microseconds waiting for the neighbours' halo stamps (k_spmv_sell: summed over the warps that
 * reach a slice reading halo blocks; divide by the number of waits), for the partial sums of p.Ap (k_update, block 0),
 * of r.z and the norm (k_direction, block 0), and the number of waits of each kind (3 more entries), and the
 * wall-clock microseconds (globaltimer) from the entry of block 0 to the
 * end of the last block summed over the launches of k_spmv_sell, k_update, k_direction (3 more entries)}.
 * reset != 0 zeroes the counters afterwards.  Zeros on the NCCL path. */
int fs_get_comm_stats(fs_context *ctx, double out[9], int reset);

/* NOTE multi-rank: fs_set_mesh and fs_destroy are collective (every rank of the communicator calls them). */

/* ---- inputs ------------------------------------------------------------ */
/* replaces the globals nu, em, thickness + initMaterialMatrices (fs.cpp:273-294) */
int fs_set_material(fs_context *ctx, double nu, double E, double thickness);
int fs_set_quirks(fs_context *ctx, int quirk_flags);
int fs_set_dof_order(fs_context *ctx, int mode);
int fs_set_assembly_mode(fs_context *ctx, int mode);
int fs_set_spmv_format(fs_context *ctx, int mode);

/* replaces mesh.read + the DirichletBoundary setup (fs.cpp:35-37, 90-120) + equation_systems.init
 * (fs.cpp:125: DOF numbering, sparsity, constraints).  The whole (replicated) mesh is passed on
 * every rank, like the reference's replicated Mesh.
 *   xyz     n_nodes*3            etype  n_elem (FS_TRI3|FS_QUAD4)
 *   eptr    n_elem+1 offsets     enodes node ids, local order
 *   bc      n_bc rows (element, side, boundary id); ids {0,20} fix u,v,w, {1,21} fix all six,
 *           {2,20,21} mark the coupling interface (fsp.cpp:68) */
int fs_set_mesh(fs_context *ctx, int64_t n_nodes, const double *xyz, int64_t n_elem,
                const int32_t *etype, const int64_t *eptr, const int32_t *enodes, int64_t n_bc,
                const int32_t *bc);

/* replaces the `forces` global filled from <mesh>_f (fs.cpp:44-67): F[n_nodes][6], already scaled */
int fs_set_nodal_loads(fs_context *ctx, const double *F);
/* replaces the coupled contribRHS (fsp.cpp:1377-1438): n interface nodes, `dims` (2|3) values each,
 * dead_axis 'x'|'y'|'z' for dims==2; all other nodal loads are zero */
int fs_set_interface_loads(fs_context *ctx, int64_t n, const int32_t *node_ids, int dims,
                           char dead_axis, const double *f);
/* scales the CURRENT device-resident load vector into the rhs: b = scale * F (constrained rows 0).
 * Used for repeated solves with a changing pressure amplitude (BASELINE config 4). */
int fs_build_rhs(fs_context *ctx, double scale);

/* ---- the hot path ------------------------------------------------------ */
/* replaces assemble_elasticity (fs.cpp:1160-1233): element kernels, Dirichlet constraint, add into
 * the block-CSR matrix, rhs.  The sparsity pattern and the element colouring are built on the
 * first call and kept.  *ms (optional) receives the device time of the values pass. */
int fs_assemble(fs_context *ctx, float *ms);
/* which kernel the last fs_assemble ran: 0 k_assemble_colored, 1 k_assemble_gather (parity block-CSR), 2 k_assemble_slice
 * (planar shells: straight into the compacted SpMV format, parity CSR on demand) */
int fs_get_assembly_path(fs_context *ctx, int *path);
/* replaces LinearImplicitSystem::solve -> KSPSolve (fs.cpp:138) for an already assembled system.
 * opts == NULL: the reference's defaults as far as they apply to CG -- rtol 1e-12, 5000 iterations, Jacobi,
 * PRECONDITIONED norm (KSPCG's default), warm start.  A solve that ends in FS_ERR_BREAKDOWN / FS_ERR_COMM leaves no
 * solution behind: fs_get_solution then fails and the next solve starts from zero. */
int fs_solve(fs_context *ctx, const fs_solve_opts *opts, fs_solve_info *info);
/* replaces build_solution_vector (fs.cpp:140-141): sols[6*node_id+var]; every rank gets the full vector */
int fs_get_solution(fs_context *ctx, double *sols);
/* stress resultants of the current solution at the element centroids, in the local axes of each element
 * (x along fs.cpp:318 / :364, z = element normal): out[6*e + k] = sigma_xx, sigma_yy, sigma_xy (membrane stress,
 * sigma = Dm B u, doc/shellelements.tex:524) and M_x, M_y, M_xy (bending moments per unit length, M = Dp B w,
 * doc/shellelements.tex:1394-1403), e = element id of fs_set_mesh.  The reference states these formulas but
 * ships no code for them; B is the element's own strain-displacement operator, so the
 * curvature sign follows it: Specht Tri-3 +d2w, DKQ Quad-4 -d2w.  Every rank receives all n_elem rows. */
int fs_recover_resultants(fs_context *ctx, double *out /* 6*n_elem */);
/* this rank's rows of the solution as PETSc holds them before build_solution_vector gathers the vector onto every
 * rank (fs.cpp:140-141): row k = k-th owned node in DOF order, node_ids[k] its mesh node id, vals[6*k+var].
 * No communication; node_ids and vals may be NULL (sizes only).  With one rank this is the whole solution. */
int fs_get_solution_owned(fs_context *ctx, int64_t *n_rows, int32_t *node_ids, double *vals);
/* equation_systems.solve() + build_solution_vector in one call with HOST buffers, as the
 * reference's coupling loop does per iteration (fsp.cpp:271-274).  F may be NULL (keep loads).
 * reassemble != 0 re-runs the values pass like the reference does on every solve.  sols may be NULL: nothing is
 * gathered and the caller reads its own rows with fs_get_solution_owned (distributed runs). */
int fs_solve_host(fs_context *ctx, const double *F, int reassemble, const fs_solve_opts *opts,
                  double *sols, fs_solve_info *info);

/* ---- coupled step (fsp.cpp:257-374) ------------------------------------ */
/* interface nodes = nodes on sides with boundary id 2, 20 or 21, ascending node id (fsp.cpp:55-71) */
int fs_interface_nodes(fs_context *ctx, int64_t *n, int32_t *node_ids /* may be NULL */);
/* forces in (dims*n_if) -> solve (K cached) -> displacement increment since the last committed
 * step out (dims*n_if), fsp.cpp:286-317 */
int fs_step(fs_context *ctx, int dims, char dead_axis, const double *forces_in, const fs_solve_opts *opts,
            double *displ_delta_out, fs_solve_info *info);
/* time step converged: preSols <- sols on the interface nodes (fsp.cpp:331-374) */
int fs_commit_step(fs_context *ctx, int dims, char dead_axis);

/* ---- parity / inspection ------------------------------------------------ */
/* sizes of the (rank-local) system: numbered nodes, 6x6 blocks, element colours (0 until a coloured pass has
 * run: the colouring is built on first use), owned node range */
int fs_get_sizes(fs_context *ctx, int64_t *n_dofnodes, int64_t *n_blocks, int64_t *n_colors,
                 int64_t *own_begin, int64_t *own_end);
/* node -> position in the DOF order (-1 for nodes no element references), n_nodes entries */
int fs_export_dof_order(fs_context *ctx, int32_t *dofnode);
/* scalar CSR of the owned rows exactly as stored on the device: rowptr (6*n_own+1, int64),
 * colidx (int32, GLOBAL dof ids; may be NULL), vals (may be NULL).  This is what `-d 1` prints
 * (fs.cpp:143-150) / -ksp_view_mat would dump. */
int fs_export_csr(fs_context *ctx, int64_t *rowptr, int32_t *colidx, double *vals);
int fs_export_rhs(fs_context *ctx, double *rhs /* 6*n_own */);
/* element matrices as the element kernel forms them, node-major, before the Dirichlet
 * constraint: Ke[e] is (6 nen)^2 doubles at out + 576*e (row stride 6 nen) */
int fs_debug_element_matrices(fs_context *ctx, double *out);
/* y = A x on the device matrix; x, y in DOF order, length 6*n_dofnodes (single rank only) */
int fs_spmv_host(fs_context *ctx, const double *x, double *y);
/* format the SpMV of the assembled matrix runs on: info = {values streamed per 6x6 block (36 parity
 * format, 14 compacted), matrix bytes streamed per SpMV (values + column ids + row/slice pointers),
 * block slots incl. padding, 36-bit union pattern of the blocks (bit 6a+b)} */
int fs_get_spmv_format(fs_context *ctx, int64_t info[4]);
/* times `reps` SpMV launches with CUDA events on the context stream (info->spmv_ms = mean) */
int fs_bench_spmv(fs_context *ctx, int reps, fs_solve_info *info);

/* measured FP64 FMA peak of the device in TFLOP/s (dependent-FMA micro-benchmark): the roofline of the
 * element kernels, which MEASURED_PEAKS.json does not carry */
int fs_bench_fp64_peak(fs_context *ctx, double *tflops);
/* north_star (a), "FP64 DMMA only if it beats the FMA pipe": the plate contraction Kp = sum_gp (Dp B)^T B of
 * fs.cpp:664-681 on a synthetic batch of n_elem Quad-4 elements, once with the production layout on the FMA pipe
 * (thread = node row) and once with mma.sync.m8n8k4.f64 (warp = element).  out = {ms per launch FMA, ms per launch
 * DMMA, max relative difference of the per-element checksums, measured DMMA peak TFLOP/s, useful TFLOP/s FMA,
 * useful TFLOP/s DMMA}.  fem_shell_b200/csrc/fs_bench.cu. */
int fs_bench_contraction(fs_context *ctx, int64_t n_elem, int reps, double out[6]);
/* micro-benchmark: microseconds per link of a chain of `links` dependent kernels (n doubles each) replayed from a CUDA
 * graph; out_us[0] = plain launches, out_us[1] = programmatic dependent launch (griddepcontrol).  fs_bench.cu. */
int fs_bench_launch_chain(fs_context *ctx, int links, int n, int reps, double out_us[2]);

/* ---- FS_PC_MLRBM (fem_shell_b200/csrc/fs_mlpc.cuh; no counterpart in the reference) ---- */
/* max_points: cap on the cells of the first lattice (default 4194304; its vectors are replicated on every rank and
 * all-reduced once per iteration).  dense_points: a lattice with at most this many cells is inverted densely
 * (default 400, maximum 512).  gamma: cycle index on the lattice levels, 1 = V, 2 = W (default); several decimal
 * digits, most significant first: digit l = visits of lattice level l+1 per visit of level l, the last digit also
 * serves every deeper level (21 = W on top, V below; 2211 = W on the two finest lattices).  The dense coarsest level is
 * always visited once per visit of the level above (it is an exact solve: a second visit computes a zero correction).
 * Same values on all ranks. */
int fs_set_ml_options(fs_context *ctx, int64_t max_points, int dense_points, int gamma);
/* *levels = number of lattice levels (0 before the first use); cells[3*l..3*l+2] = cells per axis of lattice l
 * (room for 3*14); weights[0] = estimate of lambda_max(D^-1 A) on the mesh, weights[1+l] = on lattice l (room
 * for 15; 0 for the dense level); *setup_ms = device time of the last values set-up.  Arrays may be NULL. */
int fs_get_ml_info(fs_context *ctx, int64_t *levels, int64_t *cells, double *weights, double *setup_ms);
/* several ranks: number of leading lattice levels whose cells are distributed over the ranks in slabs (halo rows
 * exchanged before each stencil application); the levels below are replicated.  0 with one rank, for lattices below
 * FS_ML_DIST_MIN_CELLS cells (environment, default 32768), or when the node blocks do not follow the slab direction. */
int fs_get_ml_dist_levels(fs_context *ctx, int64_t *n_dist);
/* number of lattice levels that iterate on the compacted copy of their stencil (18 of 36 entries per block: shells in a
 * coordinate plane; environment FS_ML_COMPACT=0 keeps the full blocks).  Results are bit-identical either way. */
int fs_get_ml_compact_levels(fs_context *ctx, int64_t *n_compact);
/* lab (environment FS_ML_PROFILE=1): FS_PC_MLRBM solves run their iterations eagerly with CUDA events between the
 * stages; ms[0..5] = accumulated time of {halo + SpMV + update, pre-smoothing + restriction to the first lattice,
 * lattice cycle, prolongation, post-smoothing SpMV + r.z, all-reduce + new direction}, ms[6..7] = the two visits of
 * the second lattice level (part of ms[2]); *iterations = how many were accumulated; reset != 0 clears them. */
int fs_get_ml_profile(fs_context *ctx, double ms[8], int64_t *iterations, int reset);
/* parity tests: copy of lattice level `level`: what = 0 the stencil (structure of arrays: entry (a,b) of the block
 * coupling cell p to its neighbour in slot s at [(6s+b)*6n + 6p + a], slot digits base 3 over the active axes,
 * offset = digit - 1), 1 the 36n pseudo-inverses of the diagonal blocks, 2 the dense inverse of the coarsest
 * level.  out = NULL: only *count is set. */
int fs_debug_ml_level(fs_context *ctx, int level, int what, double *out, int64_t capacity, int64_t *count);
/* z = M^-1 r on host vectors in dof order (6 * n_dofnodes), single rank: for parity tests of the preconditioner */
int fs_apply_mlrbm_host(fs_context *ctx, const double *r, double *z);

/* host-only (no GPU): the node-block partition plan rank `rank` of `world` derives from the replicated
 * mesh -- owned range, local nodes, local elements, halo send lists and recv segments.  Two-call
 * protocol (NULL arrays -> sizes only).  sizes = {n_dofnodes, own_begin, own_end, own_lo, n_local,
 * n_local_elems, n_send, n_peers}; peer_table rows = {rank, send_count, send_off, recv_count, recv_off}.
 * Mirrors libMesh's partitioned element loop / PETSc's VecScatter setup (fs.cpp:1197). */
int fs_partition_plan(int64_t n_nodes, int64_t n_elem, const int64_t *eptr, const int32_t *enodes, int dof_mode,
                      int rank, int world, int64_t sizes[8], int32_t *local_to_global, int32_t *loc_elems,
                      int32_t *send_idx, int64_t *peer_table);

/* host-only: the schedule of the row-gather assembly pass (fs_assemble, FS_ASM_GATHER) for a single-rank mesh given
 * in dof-node numbering -- which (element, node row) pairs each warp forms, where they add their 6x6 blocks and in
 * which conflict-free phase.  Test hook of fem_shell_b200/csrc/fs_gather_plan.cpp; needs no GPU.  Two-call protocol
 * (NULL arrays -> sizes only): sizes = {n_chunks, available (0: a block row does not fit a warp, the context would use
 * the coloured pass), n_blocks}; chunks rows = {val_off, val_count, n_phases, n_threads}; info / nodes = 32 entries of
 * 4 ints per chunk (layout: fs_gather_plan.hpp); nptr_out / nadj_out = the node-block pattern the slots refer to. */
int fs_gather_plan(int64_t n_nodes, int64_t n_elem, const int32_t *etype, const int64_t *eptr, const int32_t *enodes,
                   const uint8_t *mask /* may be NULL */, int warp_vals, int64_t sizes[3], int64_t *chunks, int32_t *info,
                   int32_t *nodes, int32_t *nptr_out, int32_t *nadj_out);

/* ---- reference file formats and generator (host side) ------------------ */
/* in-memory meshGen (src/meshgen/main_all.cpp:133-387), incl. the 6-significant-digit text
 * round trip of coordinates and load factor.  Two-call protocol: pass NULL arrays to get sizes. */
int fs_meshgen(char kind, int nx, int ny, double min_x, double min_y, double max_x, double max_y,
               const int bcids_tblr[4], double factor, int loading, int ul_lr, char dead_axis,
               int64_t *n_nodes, int64_t *n_elem, int64_t *n_bc, double *xyz, int32_t *etype,
               int64_t *eptr, int32_t *enodes, int32_t *bc, double *forces);
/* XDA reader (fs.cpp:37) and <base>_f reader (fs.cpp:44-67); two-call protocol as above */
int fs_read_xda(const char *path, int64_t *n_nodes, int64_t *n_elem, int64_t *n_enodes, int64_t *n_bc,
                double *xyz, int32_t *etype, int64_t *eptr, int32_t *enodes, int32_t *bc);
/* mesh.read() of fs.cpp:37: the format follows the extension like libMesh's -- *.msh = Gmsh MSH 2.x ASCII (3-node
 * triangles, 4-node quadrangles; 2-node lines carry the boundary ids as their physical group), *.xdr = the binary
 * (Sun XDR) twin of the XDA layout, anything else = XDA.  Same two-call protocol.  The reference holds no *.msh /
 * *.xdr fixture: these two readers follow the format descriptions (fem_shell_b200/csrc/fs_meshio.cpp). */
int fs_read_mesh(const char *path, int64_t *n_nodes, int64_t *n_elem, int64_t *n_enodes, int64_t *n_bc,
                 double *xyz, int32_t *etype, int64_t *eptr, int32_t *enodes, int32_t *bc);
int fs_write_xdr(const char *path, int64_t n_nodes, const double *xyz, int64_t n_elem,
                 const int32_t *etype, const int64_t *eptr, const int32_t *enodes, int64_t n_bc,
                 const int32_t *bc);
int fs_read_forces(const char *path, int64_t n_nodes, double *forces);
int fs_write_xda(const char *path, int64_t n_nodes, const double *xyz, int64_t n_elem,
                 const int32_t *etype, const int64_t *eptr, const int32_t *enodes, int64_t n_bc,
                 const int32_t *bc);

#ifdef __cplusplus
}
#endif
#endif /* FEMSHELL_B200_H */

// fs_assembly.cu -- sparsity pattern, element colouring and the stiffness values pass.
//
// Replaces, on the device, what the reference does per element inside assemble_elasticity
// (fs.cpp:1160-1233) through libMesh/PETSc: dof_indices (fs.cpp:1205), the element routines
// (fs.cpp:1211-1221), constrain_element_matrix_and_vector (fs.cpp:1227) and
// SparseMatrix::add_matrix -> MatSetValues(ADD_VALUES) (fs.cpp:1230), plus the sparsity pattern
// libMesh builds in equation_systems.init() (fs.cpp:125).
//
// Matrix layout in HBM ("block-CSR"): a genuine scalar CSR whose pattern is a dense 6x6 block
// per node pair sharing an element.  For owned dof-node p with deg = nptr[p+1]-nptr[p] blocks,
// scalar row 6p+a is the contiguous run  vals[36*nptr[p] + a*6*deg + 6*slot + b]  (slot = position
// of the column node in the sorted neighbour list, b = column variable).  Column indices are
// implicit (6*nadj[..]+b) and only materialised by fs_export_csr.
#include <cub/cub.cuh>

#include <algorithm>
#include <cstring>
#include <mutex>

#include "fs_context.hpp"
#include "fs_elements.cuh"
#include "fs_gather_plan.hpp"
#include "fs_sell.cuh"

namespace fs {

// ---------------------------------------------------------------------------------------------
// pattern build
// ---------------------------------------------------------------------------------------------
__global__ void k_count_candidates(const int32_t *__restrict__ conn, int nen, int64_t ne,
                                   int own_lo, int n_own, unsigned long long *cnt)
{
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= ne * nen) return;
    int p = conn[t] - own_lo;
    if (p >= 0 && p < n_own) atomicAdd(&cnt[p], (unsigned long long)nen);
}

__global__ void k_fill_candidates(const int32_t *__restrict__ conn, int nen, int64_t ne, int own_lo,
                                  int n_own, const unsigned long long *__restrict__ off,
                                  unsigned long long *fill, int32_t *cand)
{
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= ne * nen) return;
    int p = conn[t] - own_lo;
    if (p < 0 || p >= n_own) return;
    int64_t e = t / nen;
    unsigned long long slot = atomicAdd(&fill[p], (unsigned long long)nen);
    for (int l = 0; l < nen; l++) cand[off[p] + slot + l] = conn[e * nen + l];
}

// one thread per owned row: sort the candidate neighbours, drop duplicates, report the degree
__global__ void k_sort_unique(int n_own, const unsigned long long *__restrict__ off, int32_t *cand,
                              int32_t *deg)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_own) return;
    int32_t *a = cand + off[p];
    int n = (int)(off[p + 1] - off[p]);
    for (int i = 1; i < n; i++) {
        int32_t v = a[i];
        int j = i - 1;
        while (j >= 0 && a[j] > v) {
            a[j + 1] = a[j];
            j--;
        }
        a[j + 1] = v;
    }
    int u = 0;
    for (int i = 0; i < n; i++)
        if (i == 0 || a[i] != a[i - 1]) a[u++] = a[i];
    deg[p] = u;
}

__global__ void k_compact(int n_own, const unsigned long long *__restrict__ off,
                          const int32_t *__restrict__ cand, const int32_t *__restrict__ nptr,
                          int32_t *nadj)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_own) return;
    const int32_t *a = cand + off[p];
    int b = nptr[p], d = nptr[p + 1] - b;
    for (int i = 0; i < d; i++) nadj[b + i] = a[i];
}

// slot of node j in the row of node i for every (element, i, j); -1 when row i is not owned
__global__ void k_positions(const int32_t *__restrict__ conn, int nen, int64_t ne, int own_lo,
                            int n_own, const int32_t *__restrict__ nptr,
                            const int32_t *__restrict__ nadj, int32_t *pos)
{
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= ne * nen * nen) return;
    int64_t e = t / (nen * nen);
    int ij = (int)(t % (nen * nen));
    int i = ij / nen, j = ij % nen;
    int p = conn[e * nen + i] - own_lo;
    int32_t res = -1;
    if (p >= 0 && p < n_own) {
        int32_t key = conn[e * nen + j];
        int lo = nptr[p], hi = nptr[p + 1] - 1, base = lo;
        while (lo <= hi) {
            int mid = (lo + hi) >> 1;
            int32_t v = nadj[mid];
            if (v < key) lo = mid + 1;
            else if (v > key) hi = mid - 1;
            else { res = mid - base; break; }
        }
    }
    pos[t] = res;
}

// ---------------------------------------------------------------------------------------------
// element colouring (Jones-Plassmann style with hashed priorities): elements of one colour share
// no node, so their += into the matrix need no atomics.  Deterministic for a given mesh.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long elem_priority(int32_t gid)
{
    unsigned int h = (unsigned int)gid * 2654435761u;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; h *= 3266489917u; h ^= h >> 16;
    return ((unsigned long long)h << 32) | (unsigned int)(gid + 1);
}

__global__ void k_color_bid(const int32_t *__restrict__ conn, const int32_t *__restrict__ gid,
                            int nen, int64_t ne, const int32_t *__restrict__ color,
                            unsigned long long *node_best)
{
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= ne || color[e] >= 0) return;
    unsigned long long pr = elem_priority(gid[e]);
    for (int k = 0; k < nen; k++) atomicMax(&node_best[conn[e * nen + k]], pr);
}

__global__ void k_color_commit(const int32_t *__restrict__ conn, const int32_t *__restrict__ gid,
                               int nen, int64_t ne, int32_t *color,
                               const unsigned long long *__restrict__ node_best,
                               unsigned long long *node_used, unsigned int *remaining, int *overflow)
{
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= ne || color[e] >= 0) return;
    unsigned long long pr = elem_priority(gid[e]);
    unsigned long long used = 0;
    bool win = true;
    for (int k = 0; k < nen; k++) {
        int n = conn[e * nen + k];
        win = win && (node_best[n] == pr);
        used |= node_used[n];
    }
    if (!win) {
        atomicAdd(remaining, 1u);
        return;
    }
    if (~used == 0ull) {
        *overflow = 1;
        return;
    }
    int c = __ffsll((long long)~used) - 1;
    color[e] = c;
    // winners of one round never share a node, so these read-modify-writes do not race
    for (int k = 0; k < nen; k++) node_used[conn[e * nen + k]] |= (1ull << c);
}

__global__ void k_histogram(const int32_t *__restrict__ color, int64_t ne, unsigned int *hist)
{
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e < ne) atomicAdd(&hist[color[e]], 1u);
}

__global__ void k_permute_elems(const int32_t *__restrict__ conn_in, const int32_t *__restrict__ gid_in,
                                const int32_t *__restrict__ order, int nen, int64_t ne,
                                int32_t *conn_out, int32_t *gid_out)
{
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= ne) return;
    int32_t src = order[e];
    for (int k = 0; k < nen; k++) conn_out[e * nen + k] = conn_in[(int64_t)src * nen + k];
    gid_out[e] = gid_in[src];
}

__global__ void k_iota(int32_t *a, int64_t n)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) a[i] = (int32_t)i;
}

static inline unsigned int nblk(int64_t n, int bs) { return (unsigned int)((n + bs - 1) / bs); }

template <class T>
static int device_exclusive_scan(fs_context *c, const T *in, T *out, int64_t n)
{
    size_t bytes = 0;
    FS_CUDA(c, cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, n, c->stream));
    DevBuf<char> tmp;
    FS_CUDA(c, tmp.alloc(bytes));
    FS_CUDA(c, cub::DeviceScan::ExclusiveSum(tmp.p, bytes, in, out, n, c->stream));
    FS_CUDA(c, cudaStreamSynchronize(c->stream));
    return FS_OK;
}

// colour one element family (both families share node_used so colours are globally consistent)
static int color_family(fs_context *c, const int32_t *conn, const int32_t *gid, int nen, int64_t ne,
                        int32_t *color, unsigned long long *node_best, unsigned long long *node_used,
                        int64_t n_local)
{
    if (ne == 0) return FS_OK;
    DevBuf<unsigned int> rem;
    DevBuf<int> ovf;
    FS_CUDA(c, rem.alloc(1));
    FS_CUDA(c, ovf.alloc(1));
    FS_CUDA(c, cudaMemsetAsync(ovf.p, 0, sizeof(int), c->stream));
    FS_CUDA(c, cudaMemsetAsync(color, 0xff, sizeof(int32_t) * ne, c->stream));
    for (int round = 0; round < 100000; round++) {
        FS_CUDA(c, cudaMemsetAsync(node_best, 0, sizeof(unsigned long long) * n_local, c->stream));
        FS_CUDA(c, cudaMemsetAsync(rem.p, 0, sizeof(unsigned int), c->stream));
        k_color_bid<<<nblk(ne, 256), 256, 0, c->stream>>>(conn, gid, nen, ne, color, node_best);
        k_color_commit<<<nblk(ne, 256), 256, 0, c->stream>>>(conn, gid, nen, ne, color, node_best,
                                                              node_used, rem.p, ovf.p);
        unsigned int h_rem = 0;
        int h_ovf = 0;
        FS_CUDA(c, cudaMemcpyAsync(&h_rem, rem.p, sizeof h_rem, cudaMemcpyDeviceToHost, c->stream));
        FS_CUDA(c, cudaMemcpyAsync(&h_ovf, ovf.p, sizeof h_ovf, cudaMemcpyDeviceToHost, c->stream));
        FS_CUDA(c, cudaStreamSynchronize(c->stream));
        if (h_ovf) return fail(c, FS_ERR_ARG, "element colouring needs more than 64 colours");
        if (h_rem == 0) return FS_OK;
    }
    return fail(c, FS_ERR_STATE, "element colouring did not terminate");
}

static int sort_family_by_color(fs_context *c, DevBuf<int32_t> &conn, DevBuf<int32_t> &gid, int nen,
                                int64_t ne, const int32_t *color, int64_t n_colors,
                                std::vector<int64_t> &off)
{
    off.assign(n_colors + 1, 0);
    if (ne == 0) return FS_OK;
    DevBuf<int32_t> order_in, order_out, color_out, conn2, gid2;
    DevBuf<unsigned int> hist;
    FS_CUDA(c, order_in.alloc(ne));
    FS_CUDA(c, order_out.alloc(ne));
    FS_CUDA(c, color_out.alloc(ne));
    FS_CUDA(c, conn2.alloc(ne * nen));
    FS_CUDA(c, gid2.alloc(ne));
    FS_CUDA(c, hist.alloc(64));
    k_iota<<<nblk(ne, 256), 256, 0, c->stream>>>(order_in.p, ne);
    size_t bytes = 0;
    FS_CUDA(c, cub::DeviceRadixSort::SortPairs(nullptr, bytes, color, color_out.p, order_in.p,
                                               order_out.p, ne, 0, 7, c->stream));
    DevBuf<char> tmp;
    FS_CUDA(c, tmp.alloc(bytes));
    FS_CUDA(c, cub::DeviceRadixSort::SortPairs(tmp.p, bytes, color, color_out.p, order_in.p,
                                               order_out.p, ne, 0, 7, c->stream));
    k_permute_elems<<<nblk(ne, 256), 256, 0, c->stream>>>(conn.p, gid.p, order_out.p, nen, ne,
                                                           conn2.p, gid2.p);
    FS_CUDA(c, cudaMemsetAsync(hist.p, 0, 64 * sizeof(unsigned int), c->stream));
    k_histogram<<<nblk(ne, 256), 256, 0, c->stream>>>(color, ne, hist.p);
    unsigned int h[64];
    FS_CUDA(c, cudaMemcpyAsync(h, hist.p, sizeof h, cudaMemcpyDeviceToHost, c->stream));
    FS_CUDA(c, cudaStreamSynchronize(c->stream));
    for (int64_t k = 0; k < n_colors; k++) off[k + 1] = off[k] + h[k];
    std::swap(conn.p, conn2.p);
    std::swap(gid.p, gid2.p);
    return FS_OK;
}

__global__ void k_max_color(const int32_t *__restrict__ color, int64_t ne, int *mx)
{
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e < ne) atomicMax(mx, color[e]);
}

int build_pattern(fs_context *c, const std::vector<int32_t> &tri, const std::vector<int32_t> &quad,
                  const std::vector<int32_t> &tri_gid, const std::vector<int32_t> &quad_gid)
{
    const int64_t nt = c->n_tri, nq = c->n_quad;
    const int own_lo = (int)c->own_lo, n_own = (int)c->n_own;
    cudaStream_t st = c->stream;
    PhaseTimer tm("build_pattern");
    FS_CUDA(c, c->d_tri.alloc(nt * 3));
    FS_CUDA(c, c->d_quad.alloc(nq * 4));
    FS_CUDA(c, c->d_tri_gid.alloc(nt));
    FS_CUDA(c, c->d_quad_gid.alloc(nq));
    if (nt) {
        FS_CUDA(c, cudaMemcpyAsync(c->d_tri.p, tri.data(), sizeof(int32_t) * nt * 3, cudaMemcpyHostToDevice, st));
        FS_CUDA(c, cudaMemcpyAsync(c->d_tri_gid.p, tri_gid.data(), sizeof(int32_t) * nt, cudaMemcpyHostToDevice, st));
    }
    if (nq) {
        FS_CUDA(c, cudaMemcpyAsync(c->d_quad.p, quad.data(), sizeof(int32_t) * nq * 4, cudaMemcpyHostToDevice, st));
        FS_CUDA(c, cudaMemcpyAsync(c->d_quad_gid.p, quad_gid.data(), sizeof(int32_t) * nq, cudaMemcpyHostToDevice, st));
    }

    // ---- node adjacency -> nptr / nadj ----
    DevBuf<unsigned long long> cnt, off, fill;
    FS_CUDA(c, cnt.alloc(n_own + 1));
    FS_CUDA(c, off.alloc(n_own + 1));
    FS_CUDA(c, fill.alloc(n_own + 1));
    FS_CUDA(c, cudaMemsetAsync(cnt.p, 0, sizeof(unsigned long long) * (n_own + 1), st));
    FS_CUDA(c, cudaMemsetAsync(fill.p, 0, sizeof(unsigned long long) * (n_own + 1), st));
    if (nt) k_count_candidates<<<nblk(nt * 3, 256), 256, 0, st>>>(c->d_tri.p, 3, nt, own_lo, n_own, cnt.p);
    if (nq) k_count_candidates<<<nblk(nq * 4, 256), 256, 0, st>>>(c->d_quad.p, 4, nq, own_lo, n_own, cnt.p);
    int rc = device_exclusive_scan(c, cnt.p, off.p, (int64_t)n_own + 1);
    if (rc) return rc;
    unsigned long long total = 0;
    FS_CUDA(c, cudaMemcpy(&total, off.p + n_own, sizeof total, cudaMemcpyDeviceToHost));
    DevBuf<int32_t> cand, deg;
    FS_CUDA(c, cand.alloc(total));
    FS_CUDA(c, deg.alloc(n_own + 1));
    FS_CUDA(c, cudaMemsetAsync(deg.p, 0, sizeof(int32_t) * (n_own + 1), st));
    if (nt) k_fill_candidates<<<nblk(nt * 3, 256), 256, 0, st>>>(c->d_tri.p, 3, nt, own_lo, n_own, off.p, fill.p, cand.p);
    if (nq) k_fill_candidates<<<nblk(nq * 4, 256), 256, 0, st>>>(c->d_quad.p, 4, nq, own_lo, n_own, off.p, fill.p, cand.p);
    k_sort_unique<<<nblk(n_own, 128), 128, 0, st>>>(n_own, off.p, cand.p, deg.p);
    FS_CUDA(c, c->d_nptr.alloc(n_own + 1));
    rc = device_exclusive_scan(c, deg.p, c->d_nptr.p, (int64_t)n_own + 1);
    if (rc) return rc;
    int32_t nb = 0;
    FS_CUDA(c, cudaMemcpy(&nb, c->d_nptr.p + n_own, sizeof nb, cudaMemcpyDeviceToHost));
    if (nb < 0) return fail(c, FS_ERR_ARG, "more than 2^31 node blocks on one GPU");
    c->n_blocks = nb;
    FS_CUDA(c, c->d_nadj.alloc(nb));
    k_compact<<<nblk(n_own, 128), 128, 0, st>>>(n_own, off.p, cand.p, c->d_nptr.p, c->d_nadj.p);
    FS_CUDA(c, cudaStreamSynchronize(st));
    cand.release();

    tm.lap("adjacency");
    // the element colouring is only needed by the coloured scatter pass: ensure_coloring() builds it on first use
    c->n_colors = 0;
    c->colored = false;
    // ---- scatter slots ----
    FS_CUDA(c, c->d_tri_pos.alloc(nt * 9));
    FS_CUDA(c, c->d_quad_pos.alloc(nq * 16));
    if (nt) k_positions<<<nblk(nt * 9, 256), 256, 0, st>>>(c->d_tri.p, 3, nt, own_lo, n_own, c->d_nptr.p, c->d_nadj.p, c->d_tri_pos.p);
    if (nq) k_positions<<<nblk(nq * 16, 256), 256, 0, st>>>(c->d_quad.p, 4, nq, own_lo, n_own, c->d_nptr.p, c->d_nadj.p, c->d_quad_pos.p);
    FS_CUDA(c, cudaStreamSynchronize(st));
    FS_CUDA(c, cudaGetLastError());
    c->d_vals.release();   // the parity values are allocated by the first pass that writes them
    c->parity_valid = false;
    tm.lap("positions");
    rc = slice_plan_build(c);
    if (rc) return rc;
    tm.lap("slice plan");
    c->pattern_ready = true;
    return FS_OK;
}

// Element colouring + colour-sorted element arrays for k_assemble_colored (once per mesh, on first use).  The
// thread tables of the other two passes hold node ids and slots, not element indices, so they stay valid.
int ensure_coloring(fs_context *c)
{
    if (c->colored) return FS_OK;
    const int64_t nt = c->n_tri, nq = c->n_quad;
    const int own_lo = (int)c->own_lo, n_own = (int)c->n_own;
    cudaStream_t st = c->stream;
    DevBuf<unsigned long long> node_best, node_used;
    DevBuf<int32_t> tcol, qcol;
    FS_CUDA(c, node_best.alloc(c->n_local));
    FS_CUDA(c, node_used.alloc(c->n_local));
    FS_CUDA(c, tcol.alloc(nt));
    FS_CUDA(c, qcol.alloc(nq));
    FS_CUDA(c, cudaMemsetAsync(node_used.p, 0, sizeof(unsigned long long) * c->n_local, st));
    int rc = color_family(c, c->d_tri.p, c->d_tri_gid.p, 3, nt, tcol.p, node_best.p, node_used.p, c->n_local);
    if (rc) return rc;
    rc = color_family(c, c->d_quad.p, c->d_quad_gid.p, 4, nq, qcol.p, node_best.p, node_used.p, c->n_local);
    if (rc) return rc;
    DevBuf<int> mx;
    FS_CUDA(c, mx.alloc(1));
    FS_CUDA(c, cudaMemsetAsync(mx.p, 0xff, sizeof(int), st));
    if (nt) k_max_color<<<nblk(nt, 256), 256, 0, st>>>(tcol.p, nt, mx.p);
    if (nq) k_max_color<<<nblk(nq, 256), 256, 0, st>>>(qcol.p, nq, mx.p);
    int h_mx = -1;
    FS_CUDA(c, cudaMemcpyAsync(&h_mx, mx.p, sizeof h_mx, cudaMemcpyDeviceToHost, st));
    FS_CUDA(c, cudaStreamSynchronize(st));
    c->n_colors = h_mx + 1;
    rc = sort_family_by_color(c, c->d_tri, c->d_tri_gid, 3, nt, tcol.p, c->n_colors, c->tri_color_off);
    if (rc) return rc;
    rc = sort_family_by_color(c, c->d_quad, c->d_quad_gid, 4, nq, qcol.p, c->n_colors, c->quad_color_off);
    if (rc) return rc;
    if (nt) k_positions<<<nblk(nt * 9, 256), 256, 0, st>>>(c->d_tri.p, 3, nt, own_lo, n_own, c->d_nptr.p, c->d_nadj.p, c->d_tri_pos.p);
    if (nq) k_positions<<<nblk(nq * 16, 256), 256, 0, st>>>(c->d_quad.p, 4, nq, own_lo, n_own, c->d_nptr.p, c->d_nadj.p, c->d_quad_pos.p);
    FS_CUDA(c, cudaStreamSynchronize(st));
    FS_CUDA(c, cudaGetLastError());
    c->colored = true;
    return FS_OK;
}

// The material / quirk constants live in ONE __constant__ object per device (c_el); every pass that uploads and
// then reads it holds this lock until its kernels have finished, so that contexts with different materials driven
// from different host threads cannot assemble with each other's D matrices.
static std::mutex g_elconst_mutex;

int upload_element_constants(fs_context *c)
{
    const ElemConst h = make_elem_const(c->nu, c->E, c->thickness, c->quirks);
    FS_CUDA(c, upload_elem_const_tu(h, c->stream));
    if (!c->d_qgp.p) {  // Gauss-point table of the gather kernel's run-time node rows (fs_elements.cuh)
        QuadGpTab t;
        quad_gp_table(t);
        FS_CUDA(c, c->d_qgp.alloc(96));
        FS_CUDA(c, cudaMemcpyAsync(c->d_qgp.p, &t, sizeof t, cudaMemcpyHostToDevice, c->stream));
        FS_CUDA(c, cudaStreamSynchronize(c->stream));  // t is a stack object
    }
    return FS_OK;
}

// ---------------------------------------------------------------------------------------------
// values pass, coloured scatter.  Block = G groups of NEN warps; warp (g, I) handles node row I of
// 32 consecutive elements; nodal coordinates are staged once per element through shared memory.
// ---------------------------------------------------------------------------------------------
template <int NEN>
struct ScatterSink {
    double *row;          // start of scalar row 6p+0
    int L;                // 6*deg
    const int *slot;      // slot[j]
    unsigned mrow;        // Dirichlet bits of the row node
    const unsigned *mcol; // Dirichlet bits of the element's nodes
    int I;
    __device__ __forceinline__ void block(int j, const double G[6][6])
    {
        double *dst = row + 6 * slot[j];
        const unsigned mc = mcol[j];
#pragma unroll
        for (int a = 0; a < 6; a++) {
            double2 *d2 = reinterpret_cast<double2 *>(dst + (size_t)a * L);
            const bool ra = (mrow >> a) & 1u;
#pragma unroll
            for (int h = 0; h < 3; h++) {
                double v0 = G[a][2 * h], v1 = G[a][2 * h + 1];
                // fs.cpp:1227: constrained row/column -> 0, constrained diagonal -> 1 (per element)
                if (ra || ((mc >> (2 * h)) & 1u)) v0 = (ra && j == I && a == 2 * h) ? 1.0 : 0.0;
                if (ra || ((mc >> (2 * h + 1)) & 1u)) v1 = (ra && j == I && a == 2 * h + 1) ? 1.0 : 0.0;
                double2 cur = d2[h];
                cur.x += v0;
                cur.y += v1;
                d2[h] = cur;
            }
        }
    }
};

template <int NEN, int I>
__device__ __forceinline__ void scatter_row(const double *X, const int32_t *nodes, const int32_t *pos,
                                            const uint8_t *__restrict__ mask,
                                            const int32_t *__restrict__ nptr, double *vals, int own_lo)
{
    const int p = nodes[I] - own_lo;
    int slot[NEN];
    unsigned mcol[NEN];
#pragma unroll
    for (int j = 0; j < NEN; j++) {
        slot[j] = pos[I * NEN + j];
        mcol[j] = mask[nodes[j]];
    }
    if (slot[I] < 0) return;  // row not owned by this rank
    const int b0 = nptr[p], deg = nptr[p + 1] - b0;
    ScatterSink<NEN> sink;
    sink.row = vals + (size_t)36 * b0;
    sink.L = 6 * deg;
    sink.slot = slot;
    sink.mrow = mcol[I];
    sink.mcol = mcol;
    sink.I = I;
    if (NEN == 3) tri_row_blocks<I>(X, sink);
    else quad_row_blocks<I>(X, sink);
}

template <int NEN, int GROUPS>
__global__ void __launch_bounds__(32 * NEN * GROUPS)
k_assemble_colored(const int32_t *__restrict__ conn, const int32_t *__restrict__ pos, int64_t e_begin,
                   int64_t e_end, const double *__restrict__ xyz, const uint8_t *__restrict__ mask,
                   const int32_t *__restrict__ nptr, double *vals, int own_lo)
{
    __shared__ double sX[GROUPS * 32][NEN * 3 + 1];
    __shared__ int32_t sN[GROUPS * 32][NEN];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = warp / NEN, I = warp % NEN;
    const int le = grp * 32 + lane;
    const int64_t e = e_begin + (int64_t)blockIdx.x * (GROUPS * 32) + le;
    const bool live = e < e_end;
    if (live) {
        const int32_t n = conn[e * NEN + I];
        sN[le][I] = n;
        sX[le][3 * I + 0] = xyz[3 * (size_t)n + 0];
        sX[le][3 * I + 1] = xyz[3 * (size_t)n + 1];
        sX[le][3 * I + 2] = xyz[3 * (size_t)n + 2];
    }
    __syncthreads();
    if (!live) return;
    double X[NEN * 3];
    int32_t nodes[NEN];
#pragma unroll
    for (int k = 0; k < NEN * 3; k++) X[k] = sX[le][k];
#pragma unroll
    for (int k = 0; k < NEN; k++) nodes[k] = sN[le][k];
    const int32_t *ps = pos + e * (NEN * NEN);
    if (I == 0) scatter_row<NEN, 0>(X, nodes, ps, mask, nptr, vals, own_lo);
    else if (I == 1) scatter_row<NEN, 1>(X, nodes, ps, mask, nptr, vals, own_lo);
    else if (I == 2) scatter_row<NEN, 2>(X, nodes, ps, mask, nptr, vals, own_lo);
    else if (NEN == 4) scatter_row<NEN, (NEN == 4 ? 3 : 0)>(X, nodes, ps, mask, nptr, vals, own_lo);
}


// ---------------------------------------------------------------------------------------------
// values pass, row gather ("owner computes").  A WARP owns a run of block rows whose CSR values fit
// in its slice of shared memory.  Lane = one (element, node row I) incidence of those rows; it forms
// its row slice in registers (run-time I, selects only, so lanes with different I do not diverge),
// then all lanes add their j-th 6x6 block into shared memory in step j; incidences of one row that
// would meet in a slot at the same step are separated into phases by the host schedule (one phase
// on structured quads; rows are disjoint, so a step needs no atomics and only __syncwarp; the order
// of the sums is fixed by the mesh).  Finally the warp streams its rows to HBM: every CSR
// value is written exactly once, coalesced, and never read.  Warps never wait for each other.
// ---------------------------------------------------------------------------------------------
constexpr int GATHER_WARPS = 4;                       // warps (= chunks) per thread block
constexpr int GATHER_THREADS = 32 * GATHER_WARPS;
constexpr int GATHER_WARP_VALS = 2816;                // doubles of shared memory per warp (22 KB)

// rotate block j once (all lanes busy), then let the incidences of a row add it in their rounds;
// a round is 18 128-bit shared-memory read-modify-writes per lane
template <int NEN>
__device__ __forceinline__ void gather_emit(const double T[3][3], double Km[4][2][2], double Kp[4][3][3], int I,
                                            const int *slot, const unsigned *mcol, double *srow, int L,
                                            bool active, int round, int n_rounds)
{
    const unsigned mrow = I == 0 ? mcol[0] : (I == 1 ? mcol[1] : (I == 2 ? mcol[2] : mcol[3]));
#pragma unroll
    for (int j = 0; j < NEN; j++) {
        double G[6][6];
        rotate_block(T, Km[j], Kp[j], G);
        const unsigned mc = mcol[j];
        if (__any_sync(0xffffffffu, active && (mrow | mc))) {  // fs.cpp:1227, only where a Dirichlet node is involved
#pragma unroll
            for (int a = 0; a < 6; a++) {
                const bool ra = (mrow >> a) & 1u;
#pragma unroll
                for (int b = 0; b < 6; b++)
                    if (ra || ((mc >> b) & 1u)) G[a][b] = (ra && j == I && a == b) ? 1.0 : 0.0;
            }
        }
        double2 *dst = reinterpret_cast<double2 *>(srow + 6 * slot[j]);
        const int L2 = L >> 1;
        for (int r = 0; r < n_rounds; r++) {
            if (active && round == r) {
#pragma unroll
                for (int a = 0; a < 6; a++)
#pragma unroll
                    for (int h = 0; h < 3; h++) {
                        double2 cur = dst[a * L2 + h];
                        cur.x += G[a][2 * h];
                        cur.y += G[a][2 * h + 1];
                        dst[a * L2 + h] = cur;
                    }
            }
            __syncwarp();
        }
    }
}

// KINDS: 1 = the mesh has only Quad-4, 2 = only Tri-3, 3 = both (a pure mesh does not carry the other path's code)
// Inputs of a lane come from the packed thread table (32 entries per chunk, build_gather_schedule): one 16-byte
// record {meta, row info, Dirichlet bits, slots} and its node ids, both read at the kernel's first instructions
// next to the chunk record, so the only dependent global loads are the coordinates (two levels instead of four).
template <int KINDS>
__global__ void __launch_bounds__(GATHER_THREADS, 2)
k_assemble_gather(const GatherChunk *__restrict__ chunks, int n_chunks, const int4 *__restrict__ g_info,
                  const int4 *__restrict__ g_nodes, const double *__restrict__ xyz, double *__restrict__ vals,
                  const double *__restrict__ qgp)
{
    extern __shared__ __align__(128) double sv_all[];
    __shared__ __align__(16) double s_qtab[96];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ci = blockIdx.x * GATHER_WARPS + warp;
    const bool live = ci < n_chunks;
    GatherChunk ch = {0, 0, 0, 0, 0};
    int4 info = make_int4(0, 0, 0, 0), nd = make_int4(0, 0, 0, 0);
    if (live) {
        ch = chunks[ci];
        info = __ldcs(g_info + (size_t)ci * 32 + lane);
        nd = __ldcs(g_nodes + (size_t)ci * 32 + lane);
    }
    if (KINDS & 1) {
        if (threadIdx.x < 96) s_qtab[threadIdx.x] = qgp[threadIdx.x];
        __syncthreads();
    }
    if (!live) return;
    double *sv = sv_all + (size_t)warp * GATHER_WARP_VALS;
    {   // the whole slice, so that the stores do not wait for the chunk record (chunks are packed to ~92 % of it)
        double2 *z2 = reinterpret_cast<double2 *>(sv);
#pragma unroll
        for (int i = 0; i < GATHER_WARP_VALS / 64; i++) z2[i * 32 + lane] = make_double2(0.0, 0.0);
    }
    const int meta = info.x;
    const bool valid = (meta >> 8) & 1;
    const int I = meta & 3, round = (meta >> 3) & 31;
    const int is_quad = KINDS == 3 ? (meta >> 2) & 1 : (KINDS == 1);
    double Km[4][2][2], Kp[4][3][3], T[3][3];
    int slot[4];
    unsigned mcol[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        slot[k] = (info.w >> (8 * k)) & 0xff;
        mcol[k] = ((unsigned)info.z >> (8 * k)) & 0x3fu;
    }
    double *srow = sv + (info.y & 0xffff);
    const int L = 6 * ((unsigned)info.y >> 16);
    const int nodes[4] = {nd.x, nd.y, nd.z, nd.w};
    if (valid) {
        if ((KINDS & 1) && is_quad) {
            double X[12];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const size_t n = (size_t)nodes[k];
                X[3 * k] = xyz[3 * n]; X[3 * k + 1] = xyz[3 * n + 1]; X[3 * k + 2] = xyz[3 * n + 2];
            }
            QuadGeom g;
            quad_geom(X, g);
#pragma unroll
            for (int r = 0; r < 3; r++)
#pragma unroll
                for (int c2 = 0; c2 < 3; c2++) T[r][c2] = g.T[r][c2];
            quad_membrane_row_rt(g, I, Km);
            quad_plate_row_rt(g, I, s_qtab, Kp);
        } else if (KINDS & 2) {
            double X[9];
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const size_t n = (size_t)nodes[k];
                X[3 * k] = xyz[3 * n]; X[3 * k + 1] = xyz[3 * n + 1]; X[3 * k + 2] = xyz[3 * n + 2];
            }
            TriGeom g;
            tri_geom(X, g);
#pragma unroll
            for (int r = 0; r < 3; r++)
#pragma unroll
                for (int c2 = 0; c2 < 3; c2++) T[r][c2] = g.T[r][c2];
            tri_membrane_row_rt(g, I, Km);
            tri_plate_row_rt(g, I, Kp);
        }
    }
    __syncwarp();
    // a chunk holds quads in its leading lanes and triangles behind them (both only in mixed meshes)
    const bool any_quad = (KINDS & 1) && __any_sync(0xffffffffu, valid && is_quad);
    const bool any_tri = (KINDS & 2) && __any_sync(0xffffffffu, valid && !is_quad);
    if (any_quad) gather_emit<4>(T, Km, Kp, I, slot, mcol, srow, L, valid && is_quad, round, ch.n_rounds);
    if (any_tri) gather_emit<3>(T, Km, Kp, I, slot, mcol, srow, L, valid && !is_quad, round, ch.n_rounds);
    // stream the finished rows out (contiguous in the CSR value array)
    // One bulk copy (TMA engine, shared -> global) instead of 40 LDS/STG pairs per lane whose scoreboards the
    // store queue kept busy for 13 % of the warp's life: every lane orders its shared-memory writes before the
    // async proxy, one lane issues the copy and waits only until the slice has been READ (the warp then exits).
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
        const unsigned src = (unsigned)__cvta_generic_to_shared(sv);
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(vals + ch.val_off), "r"(src),
                     "r"((unsigned)ch.val_count * 8u)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
}

// host: build the thread table of the row-gather pass from the colour-sorted element arrays
int build_gather_schedule(fs_context *c)
{
    if (c->gather_ready || c->gather_unavailable) return FS_OK;
    const int64_t nt = c->n_tri, nq = c->n_quad, n_own = c->n_own;
    const int own_lo = (int)c->own_lo;
    std::vector<int32_t> tri(3 * nt), quad(4 * nq), tgid(nt), qgid(nq), nptr(n_own + 1), tpos(9 * nt), qpos(16 * nq);
    std::vector<uint8_t> mask(c->n_local);
    FS_CUDA(c, cudaMemcpy(mask.data(), c->d_mask.p, c->n_local, cudaMemcpyDeviceToHost));
    if (nt) {
        FS_CUDA(c, cudaMemcpy(tri.data(), c->d_tri.p, sizeof(int32_t) * 3 * nt, cudaMemcpyDeviceToHost));
        FS_CUDA(c, cudaMemcpy(tgid.data(), c->d_tri_gid.p, sizeof(int32_t) * nt, cudaMemcpyDeviceToHost));
        FS_CUDA(c, cudaMemcpy(tpos.data(), c->d_tri_pos.p, sizeof(int32_t) * 9 * nt, cudaMemcpyDeviceToHost));
    }
    if (nq) {
        FS_CUDA(c, cudaMemcpy(quad.data(), c->d_quad.p, sizeof(int32_t) * 4 * nq, cudaMemcpyDeviceToHost));
        FS_CUDA(c, cudaMemcpy(qgid.data(), c->d_quad_gid.p, sizeof(int32_t) * nq, cudaMemcpyDeviceToHost));
        FS_CUDA(c, cudaMemcpy(qpos.data(), c->d_quad_pos.p, sizeof(int32_t) * 16 * nq, cudaMemcpyDeviceToHost));
    }
    FS_CUDA(c, cudaMemcpy(nptr.data(), c->d_nptr.p, sizeof(int32_t) * (n_own + 1), cudaMemcpyDeviceToHost));

    GatherPlan plan;
    if (!plan_gather(n_own, own_lo, nt, tri.data(), tgid.data(), tpos.data(), nq, quad.data(), qgid.data(), qpos.data(), nptr.data(),
                     mask.data(), GATHER_WARP_VALS, plan)) {
        c->gather_unavailable = true;  // a block row exceeds a warp: leave this mesh to the coloured pass
        return FS_OK;
    }
    const std::vector<GatherChunk> &chunks = plan.chunks;
    const std::vector<int32_t> &g_info = plan.info, &g_nodes = plan.nodes;
    c->n_g_chunks = (int64_t)chunks.size();
    FS_CUDA(c, c->d_g_chunks.alloc(chunks.size()));
    FS_CUDA(c, c->d_g_info.alloc(g_info.size() / 4));
    FS_CUDA(c, c->d_g_nodes.alloc(g_nodes.size() / 4));
    FS_CUDA(c, cudaMemcpy(c->d_g_chunks.p, chunks.data(), sizeof(GatherChunk) * chunks.size(), cudaMemcpyHostToDevice));
    FS_CUDA(c, cudaMemcpy(c->d_g_info.p, g_info.data(), sizeof(int32_t) * g_info.size(), cudaMemcpyHostToDevice));
    FS_CUDA(c, cudaMemcpy(c->d_g_nodes.p, g_nodes.data(), sizeof(int32_t) * g_nodes.size(), cudaMemcpyHostToDevice));
    constexpr int smem = GATHER_WARPS * GATHER_WARP_VALS * (int)sizeof(double);
    FS_CUDA(c, cudaFuncSetAttribute(k_assemble_gather<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    FS_CUDA(c, cudaFuncSetAttribute(k_assemble_gather<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    FS_CUDA(c, cudaFuncSetAttribute(k_assemble_gather<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    c->gather_ready = true;
    return FS_OK;
}

// the parity block-CSR values (explicit zeros included) of the current material / mesh into d_vals; enqueue only
static int enqueue_parity_values(fs_context *c)
{
    cudaStream_t st = c->stream;
    if (c->d_vals.n < (size_t)36 * c->n_blocks) FS_CUDA(c, c->d_vals.alloc((size_t)36 * c->n_blocks));
    if (c->asm_mode == FS_ASM_GATHER) {
        int rc = build_gather_schedule(c);
        if (rc) return rc;
    }
    if (c->asm_mode == FS_ASM_GATHER && c->gather_ready) {
        auto kern = c->n_tri == 0 ? k_assemble_gather<1> : (c->n_quad == 0 ? k_assemble_gather<2> : k_assemble_gather<3>);
        kern<<<nblk(c->n_g_chunks, GATHER_WARPS), GATHER_THREADS, GATHER_WARPS * GATHER_WARP_VALS * sizeof(double), st>>>(
            c->d_g_chunks.p, (int)c->n_g_chunks, c->d_g_info.p, c->d_g_nodes.p, c->d_xyz.p, c->d_vals.p, c->d_qgp.p);
        FS_CUDA(c, cudaGetLastError());
        return FS_OK;
    }
    {
        int rc = ensure_coloring(c);
        if (rc) return rc;
    }
    FS_CUDA(c, cudaMemsetAsync(c->d_vals.p, 0, sizeof(double) * 36 * (size_t)c->n_blocks, st));
    constexpr int G = 2;
    for (int64_t k = 0; k < c->n_colors; k++) {
        int64_t t0 = c->tri_color_off[k], t1 = c->tri_color_off[k + 1];
        if (t1 > t0)
            k_assemble_colored<3, G><<<nblk(t1 - t0, 32 * G), 96 * G, 0, st>>>(
                c->d_tri.p, c->d_tri_pos.p, t0, t1, c->d_xyz.p, c->d_mask.p, c->d_nptr.p, c->d_vals.p, (int)c->own_lo);
        int64_t q0 = c->quad_color_off[k], q1 = c->quad_color_off[k + 1];
        if (q1 > q0)
            k_assemble_colored<4, G><<<nblk(q1 - q0, 32 * G), 128 * G, 0, st>>>(
                c->d_quad.p, c->d_quad_pos.p, q0, q1, c->d_xyz.p, c->d_mask.p, c->d_nptr.p, c->d_vals.p, (int)c->own_lo);
    }
    FS_CUDA(c, cudaGetLastError());
    return FS_OK;
}

// fs_export_csr / FS_SPMV_FULL / a non-planar union pattern need the parity values: form them from the current
// material and mesh when the last fs_assemble wrote the compacted format only
int ensure_parity_values(fs_context *c)
{
    if (c->parity_valid) return FS_OK;
    if (!c->assembled) return fail(c, FS_ERR_STATE, "matrix not assembled");
    std::lock_guard<std::mutex> lock(g_elconst_mutex);
    int rc = upload_element_constants(c);
    if (rc) return rc;
    rc = enqueue_parity_values(c);
    if (rc) return rc;
    FS_CUDA(c, cudaStreamSynchronize(c->stream));
    c->parity_valid = true;
    return FS_OK;
}

int assemble_values(fs_context *c, float *ms)
{
    cudaStream_t st = c->stream;
    std::lock_guard<std::mutex> lock(g_elconst_mutex);  // held until the stream has been synchronised
    int rc = upload_element_constants(c);
    if (rc) return rc;
    // shells in the xy plane: straight into the zero-compacted SpMV format (fs_slice_asm.cu)
    const bool direct = c->slice_ready && c->asm_mode == FS_ASM_GATHER && c->spmv_format_pref == FS_SPMV_AUTO;
    if (!direct && c->asm_mode == FS_ASM_GATHER) {  // the host plan of the row-gather pass is not part of the timed values pass
        rc = build_gather_schedule(c);
        if (rc) return rc;
    }
    FS_CUDA(c, cudaEventRecord(c->ev0, st));
    rc = direct ? assemble_slice_enqueue(c) : enqueue_parity_values(c);
    if (rc) return rc;
    FS_CUDA(c, cudaEventRecord(c->ev1, st));
    FS_CUDA(c, cudaStreamSynchronize(st));
    FS_CUDA(c, cudaGetLastError());
    if (ms) FS_CUDA(c, cudaEventElapsedTime(ms, c->ev0, c->ev1));
    c->assembled = true;
    c->minv_kind = -1;
    c->ml_values_ready = false;
    c->parity_valid = !direct;
    if (direct) {  // the iteration format IS the assembly output
        const bool was = c->sell_active && c->sell_mask == SELL_MASK_XY;
        c->sell_checked = c->sell_active = true;
        c->sell_mask = c->sell_detected = SELL_MASK_XY;
        c->sell_nz = 14;
        c->sell_kind = 0;
        if (!was && c->cg_graph_exec) { cudaGraphExecDestroy(c->cg_graph_exec); c->cg_graph_exec = nullptr; }
    } else
        c->sell_checked = false;
    return FS_OK;
}

// ---------------------------------------------------------------------------------------------
// rhs: every node loads its six values exactly once (fs.cpp:1118-1153); constrained rows -> 0
// (fs.cpp:1227).  b lives in the local vector layout (offset own_lo).
// ---------------------------------------------------------------------------------------------
__global__ void k_build_rhs(int64_t n_own, const double *__restrict__ F, const uint8_t *__restrict__ mask_own,
                            double scale, double *b_own)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= 6 * n_own) return;
    int64_t p = i / 6;
    int v = (int)(i - 6 * p);
    b_own[i] = ((mask_own[p] >> v) & 1) ? 0.0 : scale * F[i];
}

int build_rhs(fs_context *c, double scale)
{
    if (!c->loads_set) return fail(c, FS_ERR_STATE, "no loads set");
    k_build_rhs<<<nblk(6 * c->n_own, 256), 256, 0, c->stream>>>(c->n_own, c->d_F.p, c->d_mask.p + c->own_lo, scale,
                                                                 c->d_b.p + 6 * c->own_lo);
    FS_CUDA(c, cudaGetLastError());
    c->rhs_ready = true;
    return FS_OK;
}

// ---------------------------------------------------------------------------------------------
// debug: dense element matrices (node-major, unconstrained) for kernel-level parity tests
// ---------------------------------------------------------------------------------------------
template <int NEN>
struct DenseSink {
    double *K;  // (6 NEN)^2 row-major
    int I;
    __device__ __forceinline__ void block(int j, const double G[6][6])
    {
#pragma unroll
        for (int a = 0; a < 6; a++)
#pragma unroll
            for (int b = 0; b < 6; b++) K[(size_t)(6 * I + a) * (6 * NEN) + 6 * j + b] = G[a][b];
    }
};

template <int NEN>
__global__ void k_debug_elements(const int32_t *__restrict__ conn, const int32_t *__restrict__ gid,
                                 int64_t ne, const double *__restrict__ xyz, double *out)
{
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t e = t / NEN;
    int I = (int)(t % NEN);
    if (e >= ne) return;
    double X[NEN * 3];
    for (int k = 0; k < NEN; k++) {
        int n = conn[e * NEN + k];
        for (int d = 0; d < 3; d++) X[3 * k + d] = xyz[3 * (size_t)n + d];
    }
    DenseSink<NEN> sink;
    sink.K = out + (size_t)576 * gid[e];
    sink.I = I;
    if (NEN == 3) {
        if (I == 0) tri_row_blocks<0>(X, sink);
        else if (I == 1) tri_row_blocks<1>(X, sink);
        else tri_row_blocks<2>(X, sink);
    } else {
        if (I == 0) quad_row_blocks<0>(X, sink);
        else if (I == 1) quad_row_blocks<1>(X, sink);
        else if (I == 2) quad_row_blocks<2>(X, sink);
        else quad_row_blocks<3>(X, sink);
    }
}

int debug_element_matrices(fs_context *c, double *out_host)
{
    std::lock_guard<std::mutex> lock(g_elconst_mutex);
    int rc = upload_element_constants(c);
    if (rc) return rc;
    DevBuf<double> out;
    FS_CUDA(c, out.alloc((size_t)576 * c->n_elem));
    FS_CUDA(c, cudaMemsetAsync(out.p, 0, sizeof(double) * 576 * c->n_elem, c->stream));
    if (c->n_tri)
        k_debug_elements<3><<<nblk(c->n_tri * 3, 96), 96, 0, c->stream>>>(c->d_tri.p, c->d_tri_gid.p, c->n_tri, c->d_xyz.p, out.p);
    if (c->n_quad)
        k_debug_elements<4><<<nblk(c->n_quad * 4, 128), 128, 0, c->stream>>>(c->d_quad.p, c->d_quad_gid.p, c->n_quad, c->d_xyz.p, out.p);
    FS_CUDA(c, cudaMemcpyAsync(out_host, out.p, sizeof(double) * 576 * c->n_elem, cudaMemcpyDeviceToHost, c->stream));
    FS_CUDA(c, cudaStreamSynchronize(c->stream));
    FS_CUDA(c, cudaGetLastError());
    return FS_OK;
}

// ---------------------------------------------------------------------------------------------
// stress resultants at the element centroids.  The reference ships only the
// formulas (doc/shellelements.tex:524 sigma = Dm B u, :1394-1403 M = Dp B w); they are evaluated here
// with the same strain-displacement columns the stiffness kernels use (tri_bcols / quad_bcols, CST and
// bilinear membrane) on the nodal unknowns rotated into the element frame (u_loc = T u_glob, the inverse
// of fs.cpp:1094-1095).  Thread = element; the element is written by the rank owning its first node.
// out[6*gid..] = sigma_xx, sigma_yy, sigma_xy, M_x, M_y, M_xy in local element axes.
// ---------------------------------------------------------------------------------------------
template <int J>
__device__ __forceinline__ void tri_curv_add(const TriGeom &g, const TriGp &t, const double *w, double k[3])
{
    double Bc[3][3];
    tri_bcols<J>(g, t, Bc);
#pragma unroll
    for (int r = 0; r < 3; r++) k[r] += Bc[r][0] * w[3 * J] + Bc[r][1] * w[3 * J + 1] + Bc[r][2] * w[3 * J + 2];
}

template <int K>
__device__ __forceinline__ void quad_curv_add(const QuadH &h, double i00, double i01, double i10, double i11,
                                              const double *w, double k[3])
{
    double Bc[3][3];
    quad_bcols<K>(h, 0.0, 0.0, i00, i01, i10, i11, Bc);
#pragma unroll
    for (int r = 0; r < 3; r++) k[r] += Bc[r][0] * w[3 * K] + Bc[r][1] * w[3 * K + 1] + Bc[r][2] * w[3 * K + 2];
}

template <int NEN>
__global__ void __launch_bounds__(128)
k_recover_resultants(const int32_t *__restrict__ conn, const int32_t *__restrict__ gid, int64_t ne,
                     const double *__restrict__ xyz, const double *__restrict__ x, int own_lo, int n_own,
                     double *__restrict__ out)
{
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= ne) return;
    int nodes[NEN];
#pragma unroll
    for (int k = 0; k < NEN; k++) nodes[k] = conn[e * NEN + k];
    if (nodes[0] < own_lo || nodes[0] >= own_lo + n_own) return;
    double X[NEN * 3];
#pragma unroll
    for (int k = 0; k < NEN; k++)
#pragma unroll
        for (int d = 0; d < 3; d++) X[3 * k + d] = xyz[3 * (size_t)nodes[k] + d];
    double T[3][3], um[2 * NEN], wp[3 * NEN], eps[3] = {0, 0, 0}, kap[3] = {0, 0, 0};
    TriGeom tg;
    QuadGeom qg;
    if (NEN == 3) {
        tri_geom(X, tg);
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int c = 0; c < 3; c++) T[r][c] = tg.T[r][c];
    } else {
        quad_geom(X, qg);
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int c = 0; c < 3; c++) T[r][c] = qg.T[r][c];
    }
#pragma unroll
    for (int k = 0; k < NEN; k++) {
        const double *u = x + 6 * (size_t)nodes[k];
        double ul[3], tl[3];
#pragma unroll
        for (int r = 0; r < 3; r++) {
            ul[r] = T[r][0] * u[0] + T[r][1] * u[1] + T[r][2] * u[2];
            tl[r] = T[r][0] * u[3] + T[r][1] * u[4] + T[r][2] * u[5];
        }
        um[2 * k] = ul[0]; um[2 * k + 1] = ul[1];
        wp[3 * k] = ul[2]; wp[3 * k + 1] = tl[0]; wp[3 * k + 2] = tl[1];
    }
    if (NEN == 3) {
        const TriGeom &g = tg;
        const double s = 1.0 / (2.0 * g.area);
        const double px[3] = {g.y23 * s, g.y31 * s, g.y12 * s};      // fs.cpp:452-463
        const double py[3] = {-g.x23 * s, -g.x31 * s, -g.x12 * s};
#pragma unroll
        for (int k = 0; k < 3; k++) {
            eps[0] += px[k] * um[2 * k];
            eps[1] += py[k] * um[2 * k + 1];
            eps[2] += py[k] * um[2 * k] + px[k] * um[2 * k + 1];
        }
        const double C0 = g.x12 * g.x12 + g.y12 * g.y12, C1 = g.x31 * g.x31 + g.y31 * g.y31, C2 = g.x23 * g.x23 + g.y23 * g.y23;
        const double mu1 = (C0 - C1) / C2, mu2 = (C2 - C0) / C1, mu3 = (C1 - C2) / C0;  // fs.cpp:702-704
        TriGp t;
        tri_gp_terms(1.0 / 3.0, 1.0 / 3.0, mu1, mu2, mu3, t);
        double kt[3] = {0, 0, 0};
        tri_curv_add<0>(g, t, wp, kt);
        tri_curv_add<1>(g, t, wp, kt);
        tri_curv_add<2>(g, t, wp, kt);
        const double sc = 1.0 / (4.0 * g.area * g.area);  // Y of fs.cpp:578-588
        const double y20 = -2.0 * g.x23 * g.y23;
        const double y21 = (c_el.quirks & FS_Q_Y21) ? -2.0 * g.x31 * g.x31 : -2.0 * g.x31 * g.y31;
        const double y22 = -g.x23 * g.y31 - g.x31 * g.y23;
        kap[0] = (g.y23 * g.y23 * kt[0] + g.y31 * g.y31 * kt[1] + g.y23 * g.y31 * kt[2]) * sc;
        kap[1] = (g.x23 * g.x23 * kt[0] + g.x31 * g.x31 * kt[1] + g.x31 * g.x23 * kt[2]) * sc;
        kap[2] = (y20 * kt[0] + y21 * kt[1] + y22 * kt[2]) * sc;
    } else {
        const QuadGeom &g = qg;
        const double dr[4] = {-0.25, 0.25, 0.25, -0.25}, ds[4] = {-0.25, -0.25, 0.25, 0.25};  // fs.cpp:490-497 at r = s = 0
        double j00 = 0, j01 = 0, j10 = 0, j11 = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            j00 += dr[k] * g.lx[k]; j01 += dr[k] * g.ly[k];
            j10 += ds[k] * g.lx[k]; j11 += ds[k] * g.ly[k];
        }
        const double di = 1.0 / (j00 * j11 - j01 * j10);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const double px = (j11 * dr[k] - j01 * ds[k]) * di, py = (-j10 * dr[k] + j00 * ds[k]) * di;
            eps[0] += px * um[2 * k];
            eps[1] += py * um[2 * k + 1];
            eps[2] += py * um[2 * k] + px * um[2 * k + 1];
        }
        QuadH h;  // fs.cpp:613-621
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const double dx = g.dx[k], dy = g.dy[k];
            const double si = 1.0 / (dx * dx + dy * dy);
            h.a[k] = -dx * si;
            h.b[k] = 0.75 * dx * dy * si;
            h.c[k] = (0.25 * dx * dx - 0.5 * dy * dy) * si;
            h.d[k] = -dy * si;
            h.e[k] = (0.25 * dy * dy - 0.5 * dx * dx) * si;
        }
        const double p00 = 0.25 * (-g.dx[0] + g.dx[2]), p01 = 0.25 * (-g.dy[0] + g.dy[2]);  // fs.cpp:641-645 at the centre
        const double p10 = 0.25 * (-g.dx[1] + g.dx[3]), p11 = 0.25 * (-g.dy[1] + g.dy[3]);
        const double dp = 1.0 / (p00 * p11 - p01 * p10);
        const double i00 = p11 * dp, i01 = -p01 * dp, i10 = -p10 * dp, i11 = p00 * dp;
        quad_curv_add<0>(h, i00, i01, i10, i11, wp, kap);
        quad_curv_add<1>(h, i00, i01, i10, i11, wp, kap);
        quad_curv_add<2>(h, i00, i01, i10, i11, wp, kap);
        quad_curv_add<3>(h, i00, i01, i10, i11, wp, kap);
    }
    double *o = out + 6 * (size_t)gid[e];
    o[0] = c_el.dm11 * eps[0] + c_el.dm12 * eps[1];
    o[1] = c_el.dm12 * eps[0] + c_el.dm11 * eps[1];
    o[2] = c_el.dm33 * eps[2];
    o[3] = c_el.dp11 * kap[0] + c_el.dp12 * kap[1];
    o[4] = c_el.dp12 * kap[0] + c_el.dp11 * kap[1];
    o[5] = c_el.dp33 * kap[2];
}

// d_out: 6*n_elem doubles, zero-filled by the caller; d_x: solution in the local vector layout with valid halos
int recover_resultants(fs_context *c, const double *d_x, double *d_out)
{
    std::lock_guard<std::mutex> lock(g_elconst_mutex);
    int rc = upload_element_constants(c);
    if (rc) return rc;
    if (c->n_tri)
        k_recover_resultants<3><<<nblk(c->n_tri, 128), 128, 0, c->stream>>>(c->d_tri.p, c->d_tri_gid.p, c->n_tri, c->d_xyz.p, d_x,
                                                                             (int)c->own_lo, (int)c->n_own, d_out);
    if (c->n_quad)
        k_recover_resultants<4><<<nblk(c->n_quad, 128), 128, 0, c->stream>>>(c->d_quad.p, c->d_quad_gid.p, c->n_quad, c->d_xyz.p, d_x,
                                                                              (int)c->own_lo, (int)c->n_own, d_out);
    FS_CUDA(c, cudaGetLastError());
    FS_CUDA(c, cudaStreamSynchronize(c->stream));  // the constants may change once the lock is released
    return FS_OK;
}

}  // namespace fs

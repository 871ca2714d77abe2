// fs_slice_asm.cu -- values pass for shells lying in the xy plane: the element kernels write the zero-compacted
// sliced-ELL matrix the CG iteration streams (fs_sell.cuh) DIRECTLY; the parity block-CSR (explicit zeros, what
// fs_export_csr returns) is then only formed on demand.
//
// Replaces, like k_assemble_gather, the element loop of assemble_elasticity (fs.cpp:1160-1233) with its
// constrain_element_matrix_and_vector (fs.cpp:1227) and add_matrix (fs.cpp:1230) calls.
//
// Why the 22 skipped entries of every 6x6 block are EXACT zeros here: all nodes share one z, so U = B - A has
// U_z = 0 exactly, x^ = U/|U| keeps it, z^ = x^ x R has only a z component, y^ = z^ x x^ has none.  T
// (fs.cpp:378-390) is a rotation about z whose other off-diagonal entries are exact zeros, and Tt^T K Tt
// (fs.cpp:1094-1095) forms every membrane / bending coupling position from products with one of those zeros.  The
// mode is chosen from the replicated mesh (identical on every rank).
//
// Planar shells in ANY orientation take the same path in their plane frame Q (fs_context.hpp): the kernel reads node
// coordinates rotated into the frame (normal component snapped to one constant, so the argument above holds
// verbatim) and therefore forms Q~ K Q~^T; the SpMV (fs_sell.cuh, ROT) rotates x blocks in and y blocks out.  The
// Dirichlet sets of the reference fix whole translation / rotation triples (fs.cpp:90-120), which commute with Q~.
//
// Thread block = one slice of 32 consecutive owned block rows.  Thread = one (element, node row I) incidence of
// those rows, described by a 32-byte record of the slice table that is built ON THE DEVICE at fs_set_mesh time
// (no host planning, no D2H of the mesh): the incidences of a row are sorted by element id, coloured greedily
// into emit phases (two incidences of a row share a phase iff they never meet in a slot at the same local node
// index), and the slice's records are ordered (quads first, then by the incidence's rank in its row, then by row)
// so that the lanes of a warp add into DIFFERENT rows -- the shared-memory accumulator is laid out like the
// output, [slot][item][row], hence bank = row.  All lanes rotate block j, then add it in their phase; rows are
// shared by the warps of the block, so a phase ends with __syncthreads.  Summation order is a function of the
// mesh alone (no atomics).  The finished slice (dmax * 14 * 32 doubles, contiguous in the SELL array) leaves
// with ONE bulk copy (TMA engine, shared -> global): every matrix value is written once and never read.
#include <cub/cub.cuh>

#include <algorithm>
#include <mutex>

#include "fs_context.hpp"
#include "fs_elements.cuh"
#include "fs_sell.cuh"

namespace fs {

static inline unsigned int nblk(int64_t n, int bs) { return (unsigned int)std::max<int64_t>(1, (n + bs - 1) / bs); }

constexpr int SLICE_MAX_THREADS = 128;
constexpr int SLICE_MAX_ROW_INC = 24;   // incidences of one block row the plan supports (else: row-gather / coloured pass)

// ---------------------------------------------------------------------------------------------
// plan, step 1: incidences per owned row
// ---------------------------------------------------------------------------------------------
__global__ void k_sl_count(const int32_t *__restrict__ conn, int nen, int64_t ne, int own_lo, int n_own, int32_t *cnt)
{
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= ne * nen) return;
    const int p = conn[t] - own_lo;
    if (p >= 0 && p < n_own) atomicAdd(&cnt[p], 1);
}

// raw[k] = {gid, eidx << 3 | is_quad << 2 | I}
__global__ void k_sl_fill(const int32_t *__restrict__ conn, const int32_t *__restrict__ gid, int nen, int64_t ne, int own_lo, int n_own,
                          const int32_t *__restrict__ ptr, int32_t *cursor, int2 *raw)
{
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= ne * nen) return;
    const int p = conn[t] - own_lo;
    if (p < 0 || p >= n_own) return;
    const int64_t e = t / nen;
    const int I = (int)(t - e * nen);
    const int at = ptr[p] + atomicAdd(&cursor[p], 1);
    raw[at] = make_int2(gid[e], (int)(e << 3) | (nen == 4 ? 4 : 0) | I);
}

// plan, step 2 (thread = owned row): sort the row's incidences (triangles first, then by element id -- the order
// plan_gather uses), colour them into emit phases, remember each one's rank.  aux[k] = rank | phase << 8
__global__ void k_sl_rows(int n_own, const int32_t *__restrict__ ptr, int2 *raw, int32_t *aux, const int32_t *__restrict__ tri,
                          const int32_t *__restrict__ quad, int *too_many)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_own) return;
    const int b = ptr[p], n = ptr[p + 1] - b;
    if (n > SLICE_MAX_ROW_INC) {
        *too_many = 1;
        return;
    }
    int2 *a = raw + b;
    for (int i = 1; i < n; i++) {  // insertion sort: key (is_quad, gid)
        const int2 v = a[i];
        const long long kv = ((long long)((v.y >> 2) & 1) << 32) | (unsigned)v.x;
        int j = i - 1;
        while (j >= 0) {
            const long long kj = ((long long)((a[j].y >> 2) & 1) << 32) | (unsigned)a[j].x;
            if (kj <= kv) break;
            a[j + 1] = a[j];
            j--;
        }
        a[j + 1] = v;
    }
    unsigned char ph[SLICE_MAX_ROW_INC];
    for (int k = 0; k < n; k++) {
        const int qk = (a[k].y >> 2) & 1, nen = qk ? 4 : 3;
        const int32_t *ek = qk ? quad + 4 * (size_t)(a[k].y >> 3) : tri + 3 * (size_t)(a[k].y >> 3);
        unsigned used = 0;
        for (int m = 0; m < k; m++) {
            if (((a[m].y >> 2) & 1) != qk) continue;  // quads and triangles are emitted one after the other
            const int32_t *em = qk ? quad + 4 * (size_t)(a[m].y >> 3) : tri + 3 * (size_t)(a[m].y >> 3);
            bool clash = false;
            for (int j = 0; j < nen; j++) clash |= ek[j] == em[j];
            if (clash) used |= 1u << ph[m];
        }
        int c = 0;
        while ((used >> c) & 1u) c++;   // n <= 24 incidences: a free phase below 24 always exists
        ph[k] = (unsigned char)c;
        aux[b + k] = k | (c << 8);
    }
}

// plan, step 3 (block = slice): order the slice's incidences (quads first, rank-major, then row) and write the
// packed thread table.  Keys are unique small integers -> position = number of set key bits below.
//   info.x = I | is_quad << 2 | phase << 3 | valid << 8 | row_in_slice << 9
//   info.z = Dirichlet bits of the element's nodes, 8 bits each      info.w = slot of node j in the row, 8 bits each
//   nodes  = local ids of the element's nodes
// meta[s] = {emit phases of the slice, kinds present (1 quad | 2 tri)}
__global__ void __launch_bounds__(128)
k_sl_table(int n_own, const int32_t *__restrict__ ptr, const int2 *__restrict__ raw, const int32_t *__restrict__ aux,
           const int32_t *__restrict__ tri, const int32_t *__restrict__ quad, const int32_t *__restrict__ tpos,
           const int32_t *__restrict__ qpos, const uint8_t *__restrict__ mask, int4 *info, int4 *nodes, int2 *meta)
{
    constexpr int KEYS = 2 * SLICE_MAX_ROW_INC * 32;  // (tri?, rank, row)
    __shared__ unsigned bits[KEYS / 32];
    __shared__ int s_rounds, s_kinds;
    const int s = blockIdx.x;
    const int r0 = 32 * s, r1 = min(r0 + 32, n_own);
    const int i0 = ptr[r0], i1 = ptr[r1];
    for (int w = threadIdx.x; w < KEYS / 32; w += blockDim.x) bits[w] = 0;
    if (threadIdx.x == 0) s_rounds = s_kinds = 0;
    __syncthreads();
    // row of incidence k: the rows of a slice are few -> each row's thread marks its own incidences
    for (int r = r0 + (int)threadIdx.x; r < r1; r += blockDim.x)
        for (int k = ptr[r]; k < ptr[r + 1]; k++) {
            const int q = (raw[k].y >> 2) & 1, rank = aux[k] & 0xff;
            const int key = ((q ? 0 : SLICE_MAX_ROW_INC) + rank) * 32 + (r - r0);
            atomicOr(&bits[key >> 5], 1u << (key & 31));
            atomicMax(&s_rounds, (aux[k] >> 8) + 1);
            atomicOr(&s_kinds, q ? 1 : 2);
        }
    __syncthreads();
    for (int r = r0 + (int)threadIdx.x; r < r1; r += blockDim.x)
        for (int k = ptr[r]; k < ptr[r + 1]; k++) {
            const int2 v = raw[k];
            const int q = (v.y >> 2) & 1, I = v.y & 3, rank = aux[k] & 0xff, phase = aux[k] >> 8;
            const int key = ((q ? 0 : SLICE_MAX_ROW_INC) + rank) * 32 + (r - r0);
            int at = __popc(bits[key >> 5] & ((1u << (key & 31)) - 1u));
            for (int w = 0; w < (key >> 5); w++) at += __popc(bits[w]);
            const size_t e = (size_t)(v.y >> 3);
            const int nen = q ? 4 : 3;
            const int32_t *en = q ? quad + 4 * e : tri + 3 * e;
            const int32_t *ps = q ? qpos + 16 * e + 4 * I : tpos + 9 * e + 3 * I;
            unsigned slots = 0, mbits = 0;
            int nd[4] = {0, 0, 0, 0};
            for (int j = 0; j < nen; j++) {
                nd[j] = en[j];
                slots |= (unsigned)(ps[j] & 0xff) << (8 * j);
                mbits |= (unsigned)(mask[en[j]] & 0x3f) << (8 * j);
            }
            info[(size_t)i0 + at] = make_int4(I | (q << 2) | (phase << 3) | (1 << 8) | ((r - r0) << 9), 0, (int)mbits, (int)slots);
            nodes[(size_t)i0 + at] = make_int4(nd[0], nd[1], nd[2], nd[3]);
        }
    (void)i1;
    __syncthreads();
    if (threadIdx.x == 0) meta[s] = make_int2(s_rounds, s_kinds);
}

// column nodes of the sliced layout (padding slots point at the row's own node, values stay zero)
__global__ void k_sl_adj(int n_own, int own_lo, const int32_t *__restrict__ nptr, const int32_t *__restrict__ nadj,
                         const int32_t *__restrict__ sptr, int32_t *__restrict__ adj)
{
    const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (32 * s >= n_own) return;
    const int p = 32 * s + lane;
    const bool live = p < n_own;
    const int b0 = live ? nptr[p] : 0, deg = live ? nptr[p + 1] - b0 : 0;
    const int s0 = sptr[s], dmax = sptr[s + 1] - s0;
    const int self = own_lo + (live ? p : 0);
    for (int slot = 0; slot < dmax; slot++) adj[32 * (size_t)(s0 + slot) + lane] = slot < deg ? nadj[b0 + slot] : self;
}

// ---------------------------------------------------------------------------------------------
// the values pass
// ---------------------------------------------------------------------------------------------
// add the masked entries of G into the slice accumulator; sv = &acc[slot][0][row]
__device__ __forceinline__ void slice_add(double *sv, const double G[6][6])
{
#pragma unroll
    for (int a = 0; a < 6; a++)
#pragma unroll
        for (int b = 0; b < 6; b++)
            if (SELL_MASK_XY & sell_bit(a, b)) sv[32 * sell_item(SELL_MASK_XY, a, b)] += G[a][b];
}

template <int NEN>
__device__ __forceinline__ void slice_emit(const double T[3][3], double Km[4][2][2], double Kp[4][3][3], int I, const int *slot,
                                           const unsigned *mcol, double *acc_row, bool active, int phase, int n_rounds)
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
        double *dst = acc_row + (size_t)slot[j] * (32 * 14);
        for (int r = 0; r < n_rounds; r++) {
            if (active && phase == r) slice_add(dst, G);
            __syncthreads();
        }
    }
}

// KINDS: 1 = only Quad-4 in the mesh, 2 = only Tri-3, 3 = both
template <int KINDS, int MINB>
__global__ void __launch_bounds__(SLICE_MAX_THREADS, MINB)
k_assemble_slice(int n_own, const int32_t *__restrict__ inc_ptr, const int4 *__restrict__ g_info, const int4 *__restrict__ g_nodes,
                 const int2 *__restrict__ meta, const double *__restrict__ xyz, const int32_t *__restrict__ sptr,
                 double *__restrict__ sell_vals, const double *__restrict__ qgp)
{
    extern __shared__ __align__(128) double acc[];   // [dmax][14][32]
    __shared__ __align__(16) double s_qtab[96];
    const int s = blockIdx.x;
    const int r0 = 32 * s, r1 = min(r0 + 32, n_own);
    const int i0 = inc_ptr[r0], i1 = inc_ptr[r1];
    const int s0 = sptr[s], dmax = sptr[s + 1] - s0;
    const int2 mt = meta[s];
    const int n_rounds = mt.x;
    const int n_acc = dmax * (14 * 32);
    for (int i = threadIdx.x; i < n_acc / 2; i += blockDim.x) reinterpret_cast<double2 *>(acc)[i] = make_double2(0.0, 0.0);
    if (KINDS & 1)
        for (int i = threadIdx.x; i < 96; i += blockDim.x) s_qtab[i] = qgp[i];
    __syncthreads();
    for (int base = i0; base < i1; base += blockDim.x) {
        const int idx = base + (int)threadIdx.x;
        const bool valid = idx < i1;
        int4 info = make_int4(0, 0, 0, 0), nd = make_int4(0, 0, 0, 0);
        if (valid) {
            info = __ldcs(g_info + idx);
            nd = __ldcs(g_nodes + idx);
        }
        const int meta_i = info.x;
        const int I = meta_i & 3, phase = (meta_i >> 3) & 31, row = (meta_i >> 9) & 31;
        const int is_quad = KINDS == 3 ? (meta_i >> 2) & 1 : (KINDS == 1);
        double Km[4][2][2], Kp[4][3][3], T[3][3];
        int slot[4];
        unsigned mcol[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            slot[k] = (info.w >> (8 * k)) & 0xff;
            mcol[k] = ((unsigned)info.z >> (8 * k)) & 0x3fu;
        }
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
        double *acc_row = acc + row;
        // block-uniform: which element kinds this batch holds (quads sit in the leading records of a slice)
        const bool any_quad = (KINDS & 1) && __syncthreads_or(valid && is_quad);
        const bool any_tri = (KINDS & 2) && __syncthreads_or(valid && !is_quad);
        if (any_quad) slice_emit<4>(T, Km, Kp, I, slot, mcol, acc_row, valid && is_quad, phase, n_rounds);
        if (any_tri) slice_emit<3>(T, Km, Kp, I, slot, mcol, acc_row, valid && !is_quad, phase, n_rounds);
    }
    // the finished slice is contiguous in the SELL value array: one bulk copy (TMA engine, shared -> global)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0 && n_acc > 0) {
        const unsigned src = (unsigned)__cvta_generic_to_shared(acc);
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(sell_vals + (size_t)(14 * 32) * s0), "r"(src),
                     "r"((unsigned)n_acc * 8u)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
}

// diagonal (or inverse of the 6x6 diagonal block) straight from the sliced layout; entries outside the mask are zero
__global__ void k_extract_minv_sell(int n_own, int own_lo, const int32_t *__restrict__ sptr, const int32_t *__restrict__ adj,
                                    const double *__restrict__ vals, int pc, double *minv, int *bad, const __grid_constant__ PlaneQ q, int rot)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_own) return;
    const int s = p >> 5, lane = p & 31;
    const int s0 = sptr[s], dmax = sptr[s + 1] - s0;
    int slot = -1;
    for (int j = 0; j < dmax && slot < 0; j++)
        if (adj[32 * (size_t)(s0 + j) + lane] == p + own_lo) slot = j;   // the real diagonal slot precedes any padding slot
    if (slot < 0) { *bad = 1; return; }
    const double *v = vals + 32 * ((size_t)14 * (s0 + slot)) + lane;
    double M[6][12];
    for (int a = 0; a < 6; a++)
        for (int b = 0; b < 6; b++) {
            M[a][b] = (SELL_MASK_XY & sell_bit(a, b)) ? v[32 * sell_item(SELL_MASK_XY, a, b)] : 0.0;
            M[a][6 + b] = (a == b) ? 1.0 : 0.0;
        }
    if (rot) {  // the stored block is Q~ D Q~^T: back to global axes, D = Q~^T D' Q~ (3x3 pieces)
        double D[6][6];
        for (int bi = 0; bi < 2; bi++)
            for (int bj = 0; bj < 2; bj++) {
                double Tm[3][3];
                for (int i = 0; i < 3; i++)
                    for (int j = 0; j < 3; j++) Tm[i][j] = M[3 * bi + i][3 * bj] * q.m[0][j] + M[3 * bi + i][3 * bj + 1] * q.m[1][j] + M[3 * bi + i][3 * bj + 2] * q.m[2][j];
                for (int i = 0; i < 3; i++)
                    for (int j = 0; j < 3; j++) D[3 * bi + i][3 * bj + j] = q.m[0][i] * Tm[0][j] + q.m[1][i] * Tm[1][j] + q.m[2][i] * Tm[2][j];
            }
        for (int a = 0; a < 6; a++)
            for (int b = 0; b < 6; b++) M[a][b] = D[a][b];
    }
    if (pc == 1) {  // PCJacobi: a zero diagonal entry is replaced by 1 (PETSc's PCSetUp_Jacobi does the same)
        for (int a = 0; a < 6; a++) minv[6 * (size_t)p + a] = M[a][a] != 0.0 ? 1.0 / M[a][a] : 1.0;
        return;
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

// ---------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------
template <class T>
static int scan_excl(fs_context *c, const T *in, T *out, int64_t n)
{
    size_t bytes = 0;
    FS_CUDA(c, cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, n, c->stream));
    DevBuf<char> tmp;
    FS_CUDA(c, tmp.alloc(bytes));
    FS_CUDA(c, cub::DeviceScan::ExclusiveSum(tmp.p, bytes, in, out, n, c->stream));
    FS_CUDA(c, cudaStreamSynchronize(c->stream));
    return FS_OK;
}

// slice pointers / column nodes of the sliced-ELL layout for this mesh (once per mesh)
int sell_layout_build(fs_context *c)
{
    if (c->sell_layout_ready) return FS_OK;
    cudaStream_t st = c->stream;
    const int n_own = (int)c->n_own, n_slices = (n_own + 31) / 32;
    DevBuf<int32_t> dmax, dmx;
    FS_CUDA(c, dmax.alloc((size_t)n_slices + 1));
    FS_CUDA(c, dmx.alloc(1));
    FS_CUDA(c, c->d_sell_sptr.alloc((size_t)n_slices + 1));
    k_sell_dmax<<<nblk(n_slices, 8), 256, 0, st>>>(n_own, n_slices, c->d_nptr.p, dmax.p);
    int rc = scan_excl(c, dmax.p, c->d_sell_sptr.p, (int64_t)n_slices + 1);
    if (rc) return rc;
    size_t bytes = 0;
    FS_CUDA(c, cub::DeviceReduce::Max(nullptr, bytes, dmax.p, dmx.p, n_slices, st));
    DevBuf<char> tmp;
    FS_CUDA(c, tmp.alloc(bytes));
    FS_CUDA(c, cub::DeviceReduce::Max(tmp.p, bytes, dmax.p, dmx.p, n_slices, st));
    int32_t total = 0, widest = 0;
    FS_CUDA(c, cudaMemcpyAsync(&total, c->d_sell_sptr.p + n_slices, sizeof total, cudaMemcpyDeviceToHost, st));
    FS_CUDA(c, cudaMemcpyAsync(&widest, dmx.p, sizeof widest, cudaMemcpyDeviceToHost, st));
    FS_CUDA(c, cudaStreamSynchronize(st));
    c->sell_slices = n_slices;
    c->sell_slots = total;
    c->sell_dmax_max = widest;
    FS_CUDA(c, c->d_sell_adj.alloc((size_t)32 * total));
    k_sl_adj<<<nblk(n_slices, 8), 256, 0, st>>>(n_own, (int)c->own_lo, c->d_nptr.p, c->d_nadj.p, c->d_sell_sptr.p, c->d_sell_adj.p);
    if (c->world > 1) {  // peer-path SpMV: which slices read halo blocks (fs_sell.cuh)
        FS_CUDA(c, c->d_sell_hflag.alloc((size_t)n_slices));
        k_sell_halo_flags<<<nblk(n_slices, 8), 256, 0, st>>>(n_own, (int)c->own_lo, n_slices, c->d_sell_sptr.p, c->d_sell_adj.p, c->d_sell_hflag.p);
    }
    FS_CUDA(c, cudaGetLastError());
    c->sell_layout_ready = true;
    return FS_OK;
}

// device-side plan of the slice pass; leaves c->slice_ready false (no error) when this mesh does not qualify
int slice_plan_build(fs_context *c)
{
    c->slice_ready = false;
    // every node of the (replicated) mesh in one plane -> exact xy block pattern in the plane frame (header comment)
    if (!c->planar || c->n_own <= 0) return FS_OK;
    cudaStream_t st = c->stream;
    const int n_own = (int)c->n_own, own_lo = (int)c->own_lo;
    const int64_t nt = c->n_tri, nq = c->n_quad;
    int rc = sell_layout_build(c);
    if (rc) return rc;
    const size_t smem = (size_t)c->sell_dmax_max * 14 * 32 * sizeof(double);
    if (smem > 96 * 1024 || c->sell_dmax_max > 255) return FS_OK;   // two blocks per SM must fit
    DevBuf<int32_t> cnt, cursor, aux;
    DevBuf<int2> raw;
    DevBuf<int> flag;
    FS_CUDA(c, cnt.alloc((size_t)n_own + 1));
    FS_CUDA(c, cursor.alloc((size_t)n_own + 1));
    FS_CUDA(c, flag.alloc(1));
    FS_CUDA(c, c->d_sl_ptr.alloc((size_t)n_own + 1));
    FS_CUDA(c, cudaMemsetAsync(cnt.p, 0, sizeof(int32_t) * ((size_t)n_own + 1), st));
    FS_CUDA(c, cudaMemsetAsync(cursor.p, 0, sizeof(int32_t) * ((size_t)n_own + 1), st));
    FS_CUDA(c, cudaMemsetAsync(flag.p, 0, sizeof(int), st));
    if (nt) k_sl_count<<<nblk(nt * 3, 256), 256, 0, st>>>(c->d_tri.p, 3, nt, own_lo, n_own, cnt.p);
    if (nq) k_sl_count<<<nblk(nq * 4, 256), 256, 0, st>>>(c->d_quad.p, 4, nq, own_lo, n_own, cnt.p);
    rc = scan_excl(c, cnt.p, c->d_sl_ptr.p, (int64_t)n_own + 1);
    if (rc) return rc;
    int32_t total = 0;
    FS_CUDA(c, cudaMemcpy(&total, c->d_sl_ptr.p + n_own, sizeof total, cudaMemcpyDeviceToHost));
    if (total <= 0) return FS_OK;
    FS_CUDA(c, raw.alloc(total));
    FS_CUDA(c, aux.alloc(total));
    if (nt) k_sl_fill<<<nblk(nt * 3, 256), 256, 0, st>>>(c->d_tri.p, c->d_tri_gid.p, 3, nt, own_lo, n_own, c->d_sl_ptr.p, cursor.p, raw.p);
    if (nq) k_sl_fill<<<nblk(nq * 4, 256), 256, 0, st>>>(c->d_quad.p, c->d_quad_gid.p, 4, nq, own_lo, n_own, c->d_sl_ptr.p, cursor.p, raw.p);
    k_sl_rows<<<nblk(n_own, 128), 128, 0, st>>>(n_own, c->d_sl_ptr.p, raw.p, aux.p, c->d_tri.p, c->d_quad.p, flag.p);
    int h_flag = 0;
    FS_CUDA(c, cudaMemcpyAsync(&h_flag, flag.p, sizeof h_flag, cudaMemcpyDeviceToHost, st));
    FS_CUDA(c, cudaStreamSynchronize(st));
    if (h_flag) return FS_OK;   // a row with more incidences than the table supports
    const int n_slices = (int)c->sell_slices;
    FS_CUDA(c, c->d_sl_info.alloc(total));
    FS_CUDA(c, c->d_sl_nodes.alloc(total));
    FS_CUDA(c, c->d_sl_meta.alloc(n_slices));
    k_sl_table<<<n_slices, 128, 0, st>>>(n_own, c->d_sl_ptr.p, raw.p, aux.p, c->d_tri.p, c->d_quad.p, c->d_tri_pos.p, c->d_quad_pos.p,
                                         c->d_mask.p, c->d_sl_info.p, c->d_sl_nodes.p, c->d_sl_meta.p);
    FS_CUDA(c, cudaGetLastError());
    FS_CUDA(c, cudaFuncSetAttribute(k_assemble_slice<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FS_CUDA(c, cudaFuncSetAttribute(k_assemble_slice<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FS_CUDA(c, cudaFuncSetAttribute(k_assemble_slice<3, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FS_CUDA(c, cudaFuncSetAttribute(k_assemble_slice<1, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FS_CUDA(c, cudaFuncSetAttribute(k_assemble_slice<2, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const size_t need = (size_t)32 * 14 * c->sell_slots;
    if (c->d_sell_vals.n < need) FS_CUDA(c, c->d_sell_vals.alloc(need));
    FS_CUDA(c, cudaStreamSynchronize(st));
    c->slice_smem = smem;
    // pure Tri-3 meshes: 6 incidences per row = 192 per slice = three batches of 64 lanes; otherwise 128
    c->slice_threads = (nq == 0) ? 64 : SLICE_MAX_THREADS;
    c->slice_ready = true;
    return FS_OK;
}

// enqueue the slice pass (element constants must be current; caller holds the constant-memory lock)
int assemble_slice_enqueue(fs_context *c)
{
    FS_CUDA(c, upload_elem_const_tu(make_elem_const(c->nu, c->E, c->thickness, c->quirks), c->stream));  // this file's c_el
    static const int minb = getenv("FS_SLICE_MINB") ? atoi(getenv("FS_SLICE_MINB")) : 2;   // lab switch: register cap 168 instead of 255
    auto kern = c->n_tri == 0 ? (minb == 3 ? k_assemble_slice<1, 3> : k_assemble_slice<1, 2>)
                              : (c->n_quad == 0 ? (minb == 3 ? k_assemble_slice<2, 3> : k_assemble_slice<2, 2>) : k_assemble_slice<3, 2>);
    kern<<<(unsigned)c->sell_slices, c->slice_threads, c->slice_smem, c->stream>>>((int)c->n_own, c->d_sl_ptr.p, c->d_sl_info.p, c->d_sl_nodes.p,
                                                                                 c->d_sl_meta.p, c->plane_rot ? c->d_xyz_plane.p : c->d_xyz.p, c->d_sell_sptr.p, c->d_sell_vals.p,
                                                                                 c->d_qgp.p);
    FS_CUDA(c, cudaGetLastError());
    return FS_OK;
}

int extract_minv_sell(fs_context *c, int pc, int *d_bad)
{
    PlaneQ q;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) q.m[i][j] = c->plane_Q[3 * i + j];
    k_extract_minv_sell<<<nblk(c->n_own, 128), 128, 0, c->stream>>>((int)c->n_own, (int)c->own_lo, c->d_sell_sptr.p, c->d_sell_adj.p,
                                                                    c->d_sell_vals.p, pc, c->d_minv.p, d_bad, q, c->plane_rot ? 1 : 0);
    FS_CUDA(c, cudaGetLastError());
    return FS_OK;
}

}  // namespace fs
